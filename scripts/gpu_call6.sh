#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 900 python -m pytest tests/test_multigpu.py -x -q > gpurun_out/r02_pytest_multigpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_multigpu.log
tail -25 gpurun_out/r02_pytest_multigpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; echo "bench rc=$?"
tail -5 gpurun_out/r02_bench_2gpu.err
python - <<'P'
import json
try:
    j=json.loads(open("gpurun_out/r02_bench_2gpu.json").read().strip().splitlines()[-1])
    print({k:j.get(k) for k in ("value","ms_per_step","collective_ms","merge_ms","parity_check","gpu_launches")}, j["e2e"], j["roofline"].get("kernel_ms_per_step"), j["roofline"].get("frac"))
except Exception as e:
    print("bench parse failed", e)
P
