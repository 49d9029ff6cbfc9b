#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r02_bench_reference_arm.json 2> $O/r02_bench_reference_arm.err; echo "ref rc=$?"
T1=$(date +%s); echo "wall_s=$((T1-T0))" | tee -a $O/r02_bench_reference_arm.err
python - <<'P'
import json
r=json.loads(open("gpurun_out/r02_bench_reference_arm.json").read().strip().splitlines()[-1])
print("reference arm:", {k:r.get(k) for k in ("metric","value","ms_per_step","steps","warmup","sample_scale")}, r["cpu_baseline"]["cores"], r["cpu_baseline"]["thread_sweep_s"])
print("timed region s:", r["ms_per_step"]*r["steps"]/1e3)
P
