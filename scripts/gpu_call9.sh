#!/bin/bash
# 8-GPU run: multi-GPU parity tests, Flat bench at 8 (two pooling growth factors) and 4 GPUs, configs[4] (1B IVFPQ)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_multigpu.py -x -q > gpurun_out/r02_pytest_multigpu8.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_multigpu8.log
tail -4 gpurun_out/r02_pytest_multigpu8.log
run_bench() { # $1 = ngpu, $2 = tag, rest = env
  local n=$1 tag=$2; shift 2
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520 + n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02_bench_${tag}.json 2> gpurun_out/r02_bench_${tag}.err
  python - "$tag" <<'P'
import json,sys
try:
    j=json.loads(open("gpurun_out/r02_bench_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], {k:j.get(k) for k in ("value","ms_per_step","collective_ms","merge_ms")}, j["step_breakdown_ms"], "parity", j["parity_check"]["ok"], "e2e", j["e2e"]["value"])
except Exception as e:
    print(sys.argv[1], "bench parse failed", e)
P
}
run_bench 8 8gpu_g8 FB200_TC_G_SHARD=8
run_bench 8 8gpu_g16 FB200_TC_G_SHARD=16
run_bench 4 4gpu_g8 FB200_TC_G_SHARD=8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 bench_shards.py --gpus 8 --steps 5 > gpurun_out/r02_shards_8gpu.json 2> gpurun_out/r02_shards_8gpu.err; echo "shards rc=$?"
grep -E "k-means|added" gpurun_out/r02_shards_8gpu.err | head -4
python - <<'P'
import json
try:
    j=json.loads(open("gpurun_out/r02_shards_8gpu.json").read().strip().splitlines()[-1])
    print({k:j.get(k) for k in ("value","ms_per_step","parity_check")}, j["config"]["kmeans"], j["config"]["add_vec_per_s_per_gpu"], j["roofline"])
except Exception as e:
    print("parse failed", e)
P
