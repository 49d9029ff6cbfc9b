#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.log 2>&1; tail -1 $O/r02_smoke.log
python -m pytest tests -m gpu -x -q --durations=8 > $O/r02_pytest_gpu_final.log 2>&1; echo "rc=$?" >> $O/r02_pytest_gpu_final.log; tail -14 $O/r02_pytest_gpu_final.log
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02_bench_final.json 2> $O/r02_bench_final.err; echo "bench rc=$?"
/usr/bin/time -v python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r02_bench_reference_arm.json 2> $O/r02_bench_reference_arm.err; echo "ref rc=$?"
grep -E "Elapsed|Maximum resident" $O/r02_bench_reference_arm.err
python - <<'P'
import json
j=json.loads(open("gpurun_out/r02_bench_final.json").read().strip().splitlines()[-1])
print({k:j.get(k) for k in ("metric","value","ms_per_step","clocks","gpu_launches")}, j["e2e"], j["roofline"]["frac"], j["parity_check"]["ok"], j["cpu_baseline"]["value"], j["cpu_baseline"]["cores"])
for w,v in j.get("workloads",{}).items():
    print(w, v.get("value"), v.get("roofline",{}).get("frac"), v.get("cpu_baseline",{}).get("value"), v.get("recall"), v.get("error"))
r=json.loads(open("gpurun_out/r02_bench_reference_arm.json").read().strip().splitlines()[-1])
print("reference arm:", {k:r.get(k) for k in ("metric","value","ms_per_step","steps","sample_scale")}, r["cpu_baseline"]["cores"], r["cpu_baseline"]["thread_sweep_s"])
P
