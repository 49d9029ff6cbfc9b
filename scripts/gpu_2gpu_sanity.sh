#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 600 python -m pytest tests/test_multigpu.py -x -q > gpurun_out/r02_pytest_multigpu2_b.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_multigpu2_b.log
tail -4 gpurun_out/r02_pytest_multigpu2_b.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_2gpu_c.json 2> gpurun_out/r02_bench_2gpu_c.err; echo "bench rc=$?"
python - <<'P'
import json
j=json.loads(open("gpurun_out/r02_bench_2gpu_c.json").read().strip().splitlines()[-1])
print({k:j.get(k) for k in ("value","ms_per_step","collective_ms","merge_ms")}, j.get("step_breakdown_ms"), "parity", j["parity_check"]["ok"], "e2e", j["e2e"]["value"])
P
