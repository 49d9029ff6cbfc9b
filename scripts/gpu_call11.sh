#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_ivf_gpu.py tests/test_adapter_gpu.py -x -q > gpurun_out/r02_pytest_b.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_b.log
tail -15 gpurun_out/r02_pytest_b.log
./tests/adapter/_build/adapter_test > gpurun_out/r02_adapter_test.txt 2>&1; tail -8 gpurun_out/r02_adapter_test.txt
# launch list of Flat bench steps (time-only pass)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fb200 --csv --log-file gpurun_out/r02_launches_flat.csv python bench.py --steps 1 --warmup 3 --no-ivfpq --no-cpu-baseline --no-parity > gpurun_out/r02_ncu_launch_bench.log 2>&1
# full capture of the flat_tc launches of one step (7 launches; skip the 4 earlier steps)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flat_tc_kernel -s 28 -c 7 -f -o gpurun_out/r02_prof_flat_tc python bench.py --steps 1 --warmup 3 --no-ivfpq --no-cpu-baseline --no-parity > gpurun_out/r02_ncu_flat.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"tc_select_kernel|tc_rerank_kernel" -s 32 -c 8 -f -o gpurun_out/r02_prof_select python bench.py --steps 1 --warmup 3 --no-ivfpq --no-cpu-baseline --no-parity > gpurun_out/r02_ncu_select.log 2>&1
# IVF-PQ scan
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ivfpq_scan_interleaved -s 3 -c 1 -f -o gpurun_out/r02_prof_ivfpq_v3 python bench_ivf.py --index ivfpq --steps 1 --warmup 3 --recall-queries 0 > gpurun_out/r02_ncu_ivfpq_v3.log 2>&1
# k-means assignment (streaming mode) + finish + deterministic update
NQ=600000 timeout 600 ncu --set full --clock-control none -k regex:"flat_tc_kernel|tc_argmin_finish|kmeans_segment_sum" -s 2 -c 6 -f -o gpurun_out/r02_prof_kmeans python scripts/gpu_kmeans_assign.py > gpurun_out/r02_ncu_kmeans.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -6
# sanitizer over the tensor-core Flat path (hand-rolled mbarrier / TMEM hand-offs)
timeout 400 compute-sanitizer --tool memcheck python -m pytest tests/test_flat_gpu.py -x -q -k "tensor_core_path_equals_exact_path or streaming_argmin" > gpurun_out/r02_sanitizer_flat_memcheck.log 2>&1; tail -4 gpurun_out/r02_sanitizer_flat_memcheck.log
timeout 400 compute-sanitizer --tool racecheck python -m pytest tests/test_flat_gpu.py -x -q -k "streaming_argmin_equals_exact_path and 4096" > gpurun_out/r02_sanitizer_flat_racecheck.log 2>&1; tail -4 gpurun_out/r02_sanitizer_flat_racecheck.log
