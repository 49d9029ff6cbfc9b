"""k-means assignment shape: nq points against nlist centroids, k = 1, on the tcgen05 path.
Prints time, algorithmic TFLOP/s and checks the result against the exact kernel on a sample."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, faiss_b200 as fb
nq, nlist, d = int(os.environ.get("NQ", 4_000_000)), 65536, 96
g = torch.Generator(device="cuda"); g.manual_seed(3)
cent = torch.rand(nlist, d, device="cuda", generator=g)
x = torch.rand(nq, d, device="cuda", generator=g)
res = fb.StandardGpuResources()
idx = fb.GpuIndexFlatL2(res, d); idx.add(cent)
D, I = idx.search(x[:100000], 1); torch.cuda.synchronize()
t0 = time.time(); D, I = idx.search(x, 1); torch.cuda.synchronize(); t = time.time() - t0
print("assign %d x %d x %d: %.3f s, %.1f TFLOP/s algorithmic, info %s" % (nq, nlist, d, t, 2.0 * nq * nlist * d / t / 1e12, idx.lastSearchInfo()))
idx.setUseTensorCores(False)
De, Ie = idx.search(x[:20000], 1)
print("exact-kernel parity on 20000 points:", bool(torch.equal(I[:20000], Ie)), bool(torch.equal(D[:20000], De)))
