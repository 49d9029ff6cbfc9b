#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_flat_gpu.py tests/test_ivf_gpu.py -x -q > gpurun_out/r02_pytest_flat_ivf.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_flat_ivf.log
tail -4 gpurun_out/r02_pytest_flat_ivf.log
python scripts/gpu_kmeans_assign.py > gpurun_out/r02_kmeans_assign.txt 2>&1; cat gpurun_out/r02_kmeans_assign.txt | grep -v WARNING
NLIST=4096 NQ=8000000 python scripts/gpu_kmeans_assign.py > gpurun_out/r02_kmeans_assign_4096.txt 2>&1; grep -v WARNING gpurun_out/r02_kmeans_assign_4096.txt | head -4
FB200_PQ_CFG=0 python bench_ivf.py --index ivfpq --steps 5 --recall-queries 0 > gpurun_out/r02_ivfpq_c.json 2> gpurun_out/r02_ivfpq_c.err
python - <<'P'
import json
j=json.loads(open("gpurun_out/r02_ivfpq_c.json").read().strip().splitlines()[-1]); r=j["roofline"]
print("ivfpq qps %.0f ms %.2f kernel_ms %.2f frac %.3f add %.1f M/s train %.2f s" % (j["value"], j["ms_per_step"], r.get("kernel_ms_per_step",0), r.get("frac",0), j["config"]["add_vec_per_s"]/1e6, j["config"]["train_s"]))
P
