"""Round-schedule knob sweep for the Flat tensor-core search at BASELINE configs[1] (N=10M, d=128, nq=10k, k=100).
The knobs are read once per process, so every setting runs in a child process: build the index, 3 warm-up + 10 timed
searches (CUDA events, inputs resident)."""
import json
import os
import subprocess
import sys

CHILD = r"""
import os, sys, json, torch
sys.path.insert(0, os.getcwd())
import faiss_b200 as fb
g = torch.Generator(device="cuda"); g.manual_seed(1234)
N, d, nq, k = 10_000_000, 128, 10_000, 100
res = fb.StandardGpuResources()
idx = fb.GpuIndexFlatL2(res, d)
for i in range(0, N, 2_000_000):
    idx.add(torch.rand(2_000_000, d, device="cuda", generator=g))
xq = torch.rand(nq, d, device="cuda", generator=g)
for _ in range(3):
    D, I = idx.search(xq, k)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    D, I = idx.search(xq, k)
e1.record(); torch.cuda.synchronize()
print(json.dumps({"ms": e0.elapsed_time(e1) / 10, "info": idx.lastSearchInfo(), "chk": int(I.sum().item())}))
"""

settings = [
    {},
    {"FB200_TC_G_EARLY": "8"},
    {"FB200_TC_G_EARLY": "6"},
    {"FB200_TC_G_EARLY": "8", "FB200_TC_G_LATE": "8"},
    {"FB200_TC_G_LATE": "8"},
    {"FB200_TC_G_EARLY": "3", "FB200_TC_G_LATE": "4"},
    {"FB200_TC_R0": "2"},
    {"FB200_TC_R0": "3", "FB200_TC_G_EARLY": "6"},
    {"FB200_TC_G_EARLY": "8", "FB200_TC_LATE_FROM": "2048", "FB200_TC_G_LATE": "4"},
    {"FB200_TC_SELECT": "sort"},
]
out = []
for s in settings:
    env = dict(os.environ)
    env.update(s)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=300)
    line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:]
    print(json.dumps(s), line, flush=True)
    out.append({"env": s, "result": line})
json.dump(out, open("gpurun_out/r02_sweep_flat.json", "w"), indent=1)
