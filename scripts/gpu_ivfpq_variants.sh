#!/bin/bash
# IVF-PQ scan A/B on the configs[3] workload (run under gpurun): parity tests first, then bench_ivf.py with
# the kernel's tuning knobs, then one ncu --set full capture of the scan kernel.
set -x
mkdir -p gpurun_out
python -m pytest tests/test_ivf_gpu.py -x -q > gpurun_out/r02_pytest_ivf.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_ivf.log
tail -3 gpurun_out/r02_pytest_ivf.log
python bench_ivf.py --index ivfpq --steps 5 --recall-queries 0 > gpurun_out/r02_ivfpq_v2.json 2> gpurun_out/r02_ivfpq_v2.err
FB200_PQ_FADD2=1 python bench_ivf.py --index ivfpq --steps 5 --recall-queries 0 > gpurun_out/r02_ivfpq_v2_fadd2.json 2> gpurun_out/r02_ivfpq_v2_fadd2.err
FB200_PQ_GENERIC_LDS=1 python bench_ivf.py --index ivfpq --steps 5 --recall-queries 0 > gpurun_out/r02_ivfpq_v2_generic.json 2> gpurun_out/r02_ivfpq_v2_generic.err
for f in gpurun_out/r02_ivfpq_v2*.json; do python - "$f" <<'P'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=j["roofline"]
    print(sys.argv[1], "qps %.0f ms %.2f kernel_ms %.2f frac %.3f" % (j["value"], j["ms_per_step"], r.get("kernel_ms_per_step",0), r.get("frac",0)))
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
P
done
