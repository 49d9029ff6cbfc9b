#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -x -q > gpurun_out/r02_pytest_multigpu2.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_multigpu2.log
tail -12 gpurun_out/r02_pytest_multigpu2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench_shards.py --gpus 2 --steps 5 > gpurun_out/r02_shards_2gpu.json 2> gpurun_out/r02_shards_2gpu.err; echo "shards rc=$?"
grep -E "k-means|added" gpurun_out/r02_shards_2gpu.err | head -4
python - <<'P'
import json
try:
    j=json.loads(open("gpurun_out/r02_shards_2gpu.json").read().strip().splitlines()[-1])
    print({k:j.get(k) for k in ("value","ms_per_step","parity_check")}, j["config"]["kmeans"], j["roofline"].get("frac"))
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/r02_shards_2gpu.err").read()[-2000:])
P
