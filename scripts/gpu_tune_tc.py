"""schedule sweep for the flat_tc rounds: each setting in a fresh process (env read once)"""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
def run():
    import torch, time, faiss_b200 as fb
    N, d, nq, k = 10_000_000, 128, 10_000, 100
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    xb = torch.rand(N, d, device="cuda", generator=g); xq = torch.rand(nq, d, device="cuda", generator=g)
    res = fb.StandardGpuResources(); idx = fb.GpuIndexFlatL2(res, d); idx.add(xb)
    for _ in range(3): idx.search(xq, k)
    torch.cuda.synchronize(); e0 = torch.cuda.Event(True); e1 = torch.cuda.Event(True)
    import ctypes
    fb.lib.faiss_b200_kernel_timing(1)
    e0.record()
    for _ in range(5): D, I = idx.search(xq, k)
    e1.record(); torch.cuda.synchronize()
    ms = ctypes.c_double(); n = ctypes.c_int(); fb.lib.faiss_b200_kernel_timing_collect(b"flat_tc", ctypes.byref(ms), ctypes.byref(n))
    print("RESULT step %.2f ms tc %.2f ms launches/step %d fallback %s" % (e0.elapsed_time(e1) / 5, ms.value / 5, n.value // 5, idx.lastSearchInfo()), flush=True)
if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "run":  # child
        run()
    else:
        settings = [json.loads(a) for a in sys.argv[1:]] or [{}]
        for s in settings:
            env = dict(os.environ); env.update(s)
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "run"], env=env, capture_output=True, text=True, timeout=300)
            out = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
            print(json.dumps(s), out[-1] if out else ("FAIL " + r.stderr[-300:]), flush=True)
