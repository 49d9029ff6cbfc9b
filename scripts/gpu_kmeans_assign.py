"""k-means at the configs[4] shape: nq points against 65536 centroids in 96-d.
  * assignment = GpuIndexFlat k=1 (tcgen05 streaming mode): time, algorithmic TFLOP/s, parity vs the exact kernel;
  * centroid update: sort-by-assignment + segmented sum (deterministic) vs the atomic kernel;
  * a short Lloyd run end to end."""
import ctypes, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, faiss_b200 as fb
nq, nlist, d = int(os.environ.get("NQ", 4_000_000)), int(os.environ.get("NLIST", 65536)), 96
g = torch.Generator(device="cuda"); g.manual_seed(3)
cent = torch.rand(nlist, d, device="cuda", generator=g)
x = torch.rand(nq, d, device="cuda", generator=g)
res = fb.StandardGpuResources()
res.setDefaultStream(0, torch.cuda.current_stream().cuda_stream)
idx = fb.GpuIndexFlatL2(res, d); idx.add(cent)
D, I = idx.search(x[:100000], 1); torch.cuda.synchronize()
for rep in range(2):
    fb.lib.faiss_b200_kernel_timing(1)
    t0 = time.time(); D, I = idx.search(x, 1); torch.cuda.synchronize(); t = time.time() - t0
    ms, n = ctypes.c_double(), ctypes.c_int()
    fb.lib.faiss_b200_kernel_timing_collect(b"flat_tc", ctypes.byref(ms), ctypes.byref(n))
    ms2, n2 = ctypes.c_double(), ctypes.c_int()
    fb.lib.faiss_b200_kernel_timing_collect(b"tc_argmin_finish", ctypes.byref(ms2), ctypes.byref(n2))
    fb.lib.faiss_b200_kernel_timing(0)
    print("assign %d x %d x %d: %.4f s wall, %.1f TFLOP/s algorithmic (%.2f of 1440.8 sustained); flat_tc %.2f ms in %d launches, finish %.2f ms; info %s" % (
        nq, nlist, d, t, 2.0 * nq * nlist * d / t / 1e12, 2.0 * nq * nlist * d / t / 1e12 / 1440.8, ms.value, n.value, ms2.value, idx.lastSearchInfo()))
idx.setUseTensorCores(False)
De, Ie = idx.search(x[:20000], 1)
print("exact-kernel parity on 20000 points:", bool(torch.equal(I[:20000], Ie)), bool(torch.equal(D[:20000], De)))
# ---- update: deterministic (sorted) vs atomic
a = I.reshape(-1).contiguous()
for name, env in (("sorted", "0"),):
    torch.cuda.synchronize(); t0 = time.time()
    s1, c1 = fb.kmeans_accumulate(res, x, a, nlist); torch.cuda.synchronize(); t1 = time.time() - t0
    t0 = time.time()
    s2, c2 = fb.kmeans_accumulate(res, x, a, nlist); torch.cuda.synchronize(); t2 = time.time() - t0
    print("update (%s): %.2f ms first, %.2f ms second; bit-reproducible: %s; counts sum %d" % (name, t1 * 1e3, t2 * 1e3, bool(torch.equal(s1, s2)), int(c1.sum())))
ref = torch.zeros((nlist, d), dtype=torch.float64, device="cuda").index_add_(0, a, x.double())
print("update max abs err vs float64:", float((s1.double() - ref).abs().max()))
# ---- short Lloyd run
t0 = time.time()
c, obj = fb.kmeans(res, x, nlist, niter=5, seed=5, max_points_per_centroid=1 << 20)
torch.cuda.synchronize()
print("k-means 5 iterations on %d x %d -> %d centroids: %.2f s (%.3f s / iteration), objective %s" % (nq, d, nlist, time.time() - t0, (time.time() - t0) / 5, obj))
