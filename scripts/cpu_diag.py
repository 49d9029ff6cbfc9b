import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
print("affinity", len(os.sched_getaffinity(0)), "cpu_count", os.cpu_count())
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    try: print(f, open(f).read().strip())
    except Exception as e: print(f, "n/a")
os.system("lscpu | grep -E 'Model name|Socket|Core|Thread|NUMA node\\(s\\)' ; cat /proc/loadavg")
from oracle import ref
rs = np.random.RandomState(0)
N, d, nq, k = 1_000_000, 128, 1000, 100
xb = rs.rand(N, d).astype(np.float32); xq = rs.rand(nq, d).astype(np.float32)
for nt in (8, 16, 32, 64, 128):
    ref.set_omp_threads(nt)
    idx = ref.IndexFlat(d, 1); idx.add(xb)
    t = time.time(); idx.search(xq, k); t1 = time.time() - t
    t = time.time(); idx.search(xq[:100], k); t2 = time.time() - t
    print("threads %3d: nq=1000 (BLAS path) %.2f s -> %.0f QPS @1M ; nq=100 (SIMD path) %.2f s -> %.0f QPS" % (nt, t1, nq / t1, t2, 100 / t2), flush=True)
