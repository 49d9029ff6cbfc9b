#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests/test_flat_gpu.py tests/test_fullsize_gpu.py -x -q -k "not ivf" > $O/r02_pytest_c.log 2>&1; echo "rc=$?" >> $O/r02_pytest_c.log; tail -5 $O/r02_pytest_c.log
for mode in bisect sort; do
  FB200_TC_SELECT=$mode python bench.py --steps 20 --warmup 5 --no-ivfpq --no-cpu-baseline > $O/r02_bench_sel_$mode.json 2> $O/r02_bench_sel_$mode.err
  python - $mode <<'P'
import json,sys
try:
    j=json.loads(open("gpurun_out/r02_bench_sel_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], {k:j.get(k) for k in ("value","ms_per_step")}, j["step_breakdown_ms"], "parity", j["parity_check"]["ok"], "frac", j["roofline"]["frac"], j["search_info"])
except Exception as e:
    print(sys.argv[1], "bench parse failed", e); print(open("gpurun_out/r02_bench_sel_%s.err" % sys.argv[1]).read()[-1500:])
P
done
python scripts/gpu_kmeans_assign.py > $O/r02_kmeans_assign_b.txt 2>&1; grep -v WARNING $O/r02_kmeans_assign_b.txt
