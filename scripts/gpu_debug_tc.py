import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import faiss_b200 as fb
res = fb.StandardGpuResources()
for (N, d, nq, k, metric) in [(70000, 128, 300, 100, 0), (70000, 128, 300, 100, 1), (120000, 96, 130, 10, 0), (50000, 64, 64, 1, 0), (65000, 200, 40, 50, 0), (40000, 128, 520, 512, 0)]:
    g = torch.Generator(device="cuda"); g.manual_seed(N + d + k)
    xb = torch.rand(N, d, device="cuda", generator=g); xq = torch.rand(nq, d, device="cuda", generator=g)
    idx = fb.GpuIndexFlat(res, d, metric); idx.add(xb)
    D, I = idx.search(xq, k); info = idx.lastSearchInfo()
    idx.setUseTensorCores(False); De, Ie = idx.search(xq, k)
    # float64 ground truth
    S = xq.double() @ xb.double().T
    if metric == 1:
        S = (xq.double() ** 2).sum(1, keepdim=True) + (xb.double() ** 2).sum(1)[None, :] - 2 * S
        gD, gI = torch.topk(S, k, dim=1, largest=False)
    else:
        gD, gI = torch.topk(S, k, dim=1, largest=True)
    badq = (I != Ie).any(dim=1).nonzero().flatten().tolist()
    print("shape", (N, d, nq, k, metric), info, "mismatch queries TCvsExact:", len(badq), "| TC vs GT set-miss:", int((torch.sort(I,1)[0] != torch.sort(gI,1)[0]).any(1).sum()), "| exact vs GT set-miss:", int((torch.sort(Ie,1)[0] != torch.sort(gI,1)[0]).any(1).sum()), "D equal:", bool(torch.equal(D, De)))
    for q in badq[:3]:
        a, b = I[q].tolist(), Ie[q].tolist()
        pos = [j for j in range(k) if a[j] != b[j]]
        print("  q", q, "first diff pos", pos[:6], "TC ids", [a[j] for j in pos[:4]], "EX ids", [b[j] for j in pos[:4]],
              "TC D", [float(D[q, j]) for j in pos[:4]], "EX D", [float(De[q, j]) for j in pos[:4]],
              "GT", [(int(gI[q, j]), float(gD[q, j])) for j in pos[:4]])
        only_tc = set(a) - set(b); only_ex = set(b) - set(a)
        print("   only in TC:", list(only_tc)[:5], "only in exact:", list(only_ex)[:5])
