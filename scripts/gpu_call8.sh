#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench_shards.py --gpus 2 --ntotal 20000000 --nlist 16384 --steps 3 > gpurun_out/r02_shards_smoke2.json 2> gpurun_out/r02_shards_smoke2.err; echo "shards rc=$?"
tail -4 gpurun_out/r02_shards_smoke2.err
python - <<'P'
import json
try:
    j=json.loads(open("gpurun_out/r02_shards_smoke2.json").read().strip().splitlines()[-1])
    print({k:j.get(k) for k in ("value","ms_per_step","parity_check")}, j["config"]["kmeans"], j["roofline"].get("frac"))
except Exception as e:
    print("parse failed", e)
P
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_2gpu_b.json 2> gpurun_out/r02_bench_2gpu_b.err; echo "bench rc=$?"
python - <<'P'
import json
try:
    j=json.loads(open("gpurun_out/r02_bench_2gpu_b.json").read().strip().splitlines()[-1])
    print({k:j.get(k) for k in ("value","ms_per_step","collective_ms","merge_ms","step_breakdown_ms")}, j["parity_check"]["ok"])
except Exception as e:
    print("bench parse failed", e)
P
