"""Wider Flat shapes on the tensor-core path: K-split kernel (128 < d <= 256), k up to 2048, fp16 storage.
Times the tcgen05 path against the exact SIMT kernel on the same index (CUDA events, inputs resident)."""
import json
import sys
import time

import numpy as np
import torch

import faiss_b200 as fb


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    res = fb.StandardGpuResources()
    out = []
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    shapes = [
        # N, d, nq, k, fp16
        (2_000_000, 256, 4096, 100, False),
        (2_000_000, 192, 4096, 100, False),
        (2_000_000, 128, 2048, 1024, False),
        (2_000_000, 128, 2048, 2048, False),
        (2_000_000, 128, 4096, 512, False),
        (10_000_000, 128, 10000, 100, True),
        (1_000_000, 256, 65536, 1, False),
    ]
    only = sys.argv[1:] and [int(a) for a in sys.argv[1:]]
    for si, (N, d, nq, k, f16) in enumerate(shapes):
        if only and si not in only:
            continue
        xb = torch.rand(N, d, device="cuda", generator=g)
        xq = torch.rand(nq, d, device="cuda", generator=g)
        idx = fb.GpuIndexFlat(res, d, fb.METRIC_L2, use_float16=f16)
        idx.add(xb)
        del xb
        D, I = idx.search(xq, k)
        info = idx.lastSearchInfo()
        t_tc = timed(lambda: idx.search(xq, k))
        idx.setUseTensorCores(False)
        nqe = min(nq, 512)  # the exact kernel is slow at these sizes: time a slice, report per query
        De, Ie = idx.search(xq[:nqe], k)
        t_ex = timed(lambda: idx.search(xq[:nqe], k), reps=1)
        same = bool(torch.equal(I[:nqe], Ie) and torch.equal(D[:nqe], De))
        rec = {
            "N": N, "d": d, "nq": nq, "k": k, "fp16_storage": f16, "tensor_cores": info["tensor_cores"],
            "fallback_queries": info["fallback_queries"], "tc_ms": round(t_tc, 3), "tc_qps": round(nq / t_tc * 1e3, 1),
            "tc_tflops": round(2.0 * nq * N * d / t_tc / 1e9, 1),
            "exact_ms_per_%d" % nqe: round(t_ex, 3), "exact_qps": round(nqe / t_ex * 1e3, 1), "identical_to_exact": same,
        }
        print(json.dumps(rec), flush=True)
        out.append(rec)
        del idx
        torch.cuda.empty_cache()
    json.dump(out, open("gpurun_out/r02_widepath.json", "w"), indent=1)


if __name__ == "__main__":
    main()
