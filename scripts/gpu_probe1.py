"""First on-GPU probe: each section runs in its own process under a timeout (a tcgen05 deadlock
must not hang the box).  Usage: python scripts/gpu_probe1.py [section]"""
import os, sys, subprocess, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

SECTIONS = ["exact", "tcdebug", "tcsearch", "ivf", "kmeans"]

def sec_exact():
    import torch, faiss_b200 as fb
    from oracle import oracle_np as o
    res = fb.StandardGpuResources()
    rs = np.random.RandomState(1)
    for (N, d, nq, k, metric) in [(3000, 32, 40, 10, 1), (5000, 128, 100, 100, 1), (4000, 33, 17, 1, 1), (3000, 64, 30, 300, 0), (2500, 16, 9, 2048, 1)]:
        xb = rs.rand(N, d).astype(np.float32); xq = rs.rand(nq, d).astype(np.float32)
        idx = fb.GpuIndexFlat(res, d, metric, use_tensor_cores=False); idx.add(xb)
        t = time.time(); D, I = idx.search(xq, k); t = time.time() - t
        Dr, Ir = o.knn_flat(xq, xb, k, metric)
        print("exact N=%d d=%d nq=%d k=%d m=%d: id match %.4f, max rel err %.2e, %.1f ms" % (N, d, nq, k, metric, (I == Ir).mean(), np.abs(D - Dr).max() / max(1e-9, np.abs(Dr[Ir>=0]).max()), t * 1e3), flush=True)
    # integer regime: ids must match exactly
    xb = np.floor(rs.rand(20000, 64) * 16).astype(np.float32); xq = np.floor(rs.rand(50, 64) * 16).astype(np.float32)
    idx = fb.GpuIndexFlatL2(res, 64, use_tensor_cores=False); idx.add(xb)
    D, I = idx.search(xq, 50); Dr, Ir = o.knn_flat(xq, xb, 50, 1)
    print("exact integer regime: ids equal", (I == Ir).all(), "D equal", (D == Dr).all(), flush=True)

def sec_tcdebug():
    import torch, faiss_b200 as fb
    res = fb.StandardGpuResources()
    torch.manual_seed(0)
    for (nq, N, dpad) in [(128, 256, 64), (200, 1000, 128), (128, 128 * 40, 128), (300, 5000, 256)]:
        Q = (torch.randn(nq, dpad, device="cuda")).half(); Y = (torch.randn(N, dpad, device="cuda")).half()
        S = fb.flat_tc_scores_debug(res, Q, Y); torch.cuda.synchronize()
        ref = Q.float() @ Y.float().T
        err = (S[:, :N] - ref).abs().max().item()
        print("tcdebug nq=%d N=%d dpad=%d: max abs err %.3e (ref max %.1f)" % (nq, N, dpad, err, ref.abs().max().item()), flush=True)

def sec_tcsearch():
    import torch, faiss_b200 as fb
    res = fb.StandardGpuResources()
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    for (N, d, nq, k, metric) in [(100000, 128, 1000, 10, 1), (200000, 128, 500, 100, 1), (150000, 96, 300, 100, 0), (1000000, 128, 2000, 100, 1)]:
        xb = torch.rand(N, d, device="cuda", generator=g); xq = torch.rand(nq, d, device="cuda", generator=g)
        idx = fb.GpuIndexFlat(res, d, metric); idx.add(xb)
        torch.cuda.synchronize(); t = time.time(); D, I = idx.search(xq, k); torch.cuda.synchronize(); t1 = time.time() - t
        info = idx.lastSearchInfo()
        t = time.time(); D, I = idx.search(xq, k); torch.cuda.synchronize(); t2 = time.time() - t
        idx.setUseTensorCores(False)
        t = time.time(); De, Ie = idx.search(xq, k); torch.cuda.synchronize(); t3 = time.time() - t
        print("tcsearch N=%d d=%d nq=%d k=%d m=%d: ids equal %s (%.5f) D equal %s | tc %.1f/%.1f ms exact %.1f ms info %s" % (
            N, d, nq, k, metric, bool((I == Ie).all()), (I == Ie).float().mean().item(), bool((D == De).all()), t1 * 1e3, t2 * 1e3, t3 * 1e3, info), flush=True)

def sec_ivf():
    import torch, faiss_b200 as fb
    from oracle import oracle_np as o
    res = fb.StandardGpuResources()
    rs = np.random.RandomState(3)
    N, d, nlist, M, nq, k = 20000, 32, 64, 8, 50, 10
    xb = rs.rand(N, d).astype(np.float32); xq = rs.rand(nq, d).astype(np.float32)
    for metric in (1, 0):
        ivf = fb.GpuIndexIVFFlat(res, d, nlist, metric); ivf.train(xb); ivf.add(xb); ivf.nprobe = 8
        D, I = ivf.search(xq, k)
        cent = ivf.getCoarseCentroids()
        lens = [ivf.getListLength(l) for l in range(nlist)]
        lv = [ivf.getListVectorData(l).view(np.float32) for l in range(nlist)]; li = [ivf.getListIndices(l) for l in range(nlist)]
        Dr, Ir = o.ivfflat_search(xq, k, 8, cent, lv, li, metric)
        print("ivfflat m=%d: sum lens %d, id match %.4f, max abs D err %.2e" % (metric, sum(lens), (I == Ir).mean(), np.abs(D - Dr).max()), flush=True)
        pq = fb.GpuIndexIVFPQ(res, d, nlist, M, 8, metric); pq.train(xb); pq.add(xb); pq.nprobe = 8
        D, I = pq.search(xq, k)
        cent = pq.getCoarseCentroids(); pqc = pq.getPQCentroids()
        lc = [pq.getListVectorData(l) for l in range(nlist)]; li = [pq.getListIndices(l) for l in range(nlist)]
        Dr, Ir = o.ivfpq_search(xq, k, 8, cent, pqc, lc, li, metric)
        print("ivfpq m=%d: ntotal %d, id match %.4f, max abs D err %.2e" % (metric, pq.ntotal, (I == Ir).mean(), np.abs(D - Dr).max()), flush=True)
        # encode parity
        a = o.ivf_assign(xb[:2000], cent, metric)
        codes = o.pq_encode(xb[:2000] - cent[a], pqc)
        got = {}
        for l in range(nlist):
            for c, i in zip(lc[l].reshape(-1, M), li[l]):
                if i < 2000: got[int(i)] = c
        mism = sum(int((got[i] != codes[i]).any()) for i in range(2000))
        print("ivfpq encode mismatching vectors: %d / 2000" % mism, flush=True)

def sec_kmeans():
    import faiss_b200 as fb
    from oracle import oracle_np as o
    res = fb.StandardGpuResources()
    rs = np.random.RandomState(7)
    x = rs.rand(20000, 16).astype(np.float32)
    t = time.time(); c, obj = fb.kmeans(res, x, 50, niter=10, seed=123); t = time.time() - t
    co, objo = o.kmeans(x, 50, niter=10, seed=123)
    print("kmeans obj gpu", obj[[0, -1]], "oracle", objo[[0, -1]], "centroid max diff %.3e, %.1f ms" % (np.abs(c - co).max(), t * 1e3), flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1:
        globals()["sec_" + sys.argv[1]]()
    else:
        for s in SECTIONS:
            print("=== section", s, flush=True)
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), s], timeout=240)
                print("=== section", s, "exit", r.returncode, flush=True)
            except subprocess.TimeoutExpired:
                print("=== section", s, "TIMEOUT", flush=True)
