#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_ivf_gpu.py -x -q > gpurun_out/r02_pytest_ivf.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_ivf.log
tail -3 gpurun_out/r02_pytest_ivf.log
for cfg in 0 3; do
  FB200_PQ_CFG=$cfg python bench_ivf.py --index ivfpq --steps 5 --recall-queries 0 > gpurun_out/r02_ivfpq_cfg$cfg.json 2> gpurun_out/r02_ivfpq_cfg$cfg.err
  FB200_PQ_CFG=$cfg python bench_ivf.py --index ivfpq --n 20000000 --nlist 10240 --steps 5 --recall-queries 0 > gpurun_out/r02_ivfpq_short_cfg$cfg.json 2> gpurun_out/r02_ivfpq_short_cfg$cfg.err
done
for f in gpurun_out/r02_ivfpq_cfg[03].json gpurun_out/r02_ivfpq_short_cfg[03].json; do python - "$f" <<'P'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=j["roofline"]
    print(sys.argv[1], "qps %.0f ms %.2f kernel_ms %.2f frac %.3f" % (j["value"], j["ms_per_step"], r.get("kernel_ms_per_step",0), r.get("frac",0)), j["clocks"].get("sm_mhz"))
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
P
done
