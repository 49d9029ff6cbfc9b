#!/bin/bash
# configs[4]: 1 B x 96-d IVFPQ over 8 shards (k-means sharded in C++, one all-reduce per iteration), then the Flat bench at 8
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 bench_shards.py --gpus 8 --steps 5 > gpurun_out/r02_shards_8gpu_b.json 2> gpurun_out/r02_shards_8gpu_b.err; echo "shards rc=$?"
grep -E "k-means|added" gpurun_out/r02_shards_8gpu_b.err | tr '[' '\n' | grep -E "rank 0" | head -4
python - <<'P'
import json
try:
    j=json.loads(open("gpurun_out/r02_shards_8gpu_b.json").read().strip().splitlines()[-1])
    print({k:j.get(k) for k in ("value","ms_per_step","parity_check")}, j["config"].get("kmeans"), j["config"].get("add_vec_per_s_per_gpu"), j["roofline"])
except Exception as e:
    print("parse failed", e)
P
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_8gpu_c.json 2> gpurun_out/r02_bench_8gpu_c.err; echo "bench rc=$?"
python - <<'P'
import json
try:
    j=json.loads(open("gpurun_out/r02_bench_8gpu_c.json").read().strip().splitlines()[-1])
    print({k:j.get(k) for k in ("value","ms_per_step","collective_ms","merge_ms")}, j.get("step_breakdown_ms"), "parity", j["parity_check"]["ok"], "e2e", j["e2e"]["value"])
except Exception as e:
    print("parse failed", e)
P
