#!/bin/bash
# ncu captures summarised ON the box (the .ncu-rep files are too large to bring back: 64 MiB cap on gpurun_out/)
mkdir -p gpurun_out
O=gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fb200 --csv --log-file $O/r02_launches_flat.csv python bench.py --steps 1 --warmup 3 --no-ivfpq --no-cpu-baseline --no-parity > $O/r02_ncu_launch_bench.log 2>&1
python scripts/launch_summary.py $O/r02_launches_flat.csv > $O/r02_launches_flat_step.txt 2>&1; tail -12 $O/r02_launches_flat_step.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flat_tc_kernel -s 28 -c 7 -f -o $O/prof_flat python bench.py --steps 1 --warmup 3 --no-ivfpq --no-cpu-baseline --no-parity > $O/r02_ncu_flat.log 2>&1
python scripts/ncu_summary.py $O/prof_flat.ncu-rep > $O/r02_ncu_flat_tc_summary.txt 2>&1
python scripts/ncu_traffic2.py $O/prof_flat.ncu-rep $O/flat_tc_traffic.json "flat_tc_kernel launches of one bench.py step, N=10M d=128 nq=10k k=100" > /dev/null 2>&1
rm -f $O/prof_flat.ncu-rep
timeout 600 ncu --set full --clock-control none -k regex:"tc_select_kernel|tc_rerank_kernel" -s 32 -c 8 -f -o $O/prof_sel python bench.py --steps 1 --warmup 3 --no-ivfpq --no-cpu-baseline --no-parity > $O/r02_ncu_select.log 2>&1
python scripts/ncu_summary.py $O/prof_sel.ncu-rep > $O/r02_ncu_select_rerank_summary.txt 2>&1; rm -f $O/prof_sel.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ivfpq_scan_interleaved -s 3 -c 1 -f -o $O/r02_prof_ivfpq_v3 python bench_ivf.py --index ivfpq --steps 1 --warmup 3 --recall-queries 0 > $O/r02_ncu_ivfpq_v3.log 2>&1
python scripts/ncu_summary.py $O/r02_prof_ivfpq_v3.ncu-rep > $O/r02_ncu_ivfpq_v3_summary.txt 2>&1
python scripts/ncu_traffic2.py $O/r02_prof_ivfpq_v3.ncu-rep $O/ivfpq_scan_traffic.json "ivfpq_scan_interleaved_kernel, one bench_ivf.py step, N=100M nlist=4096 M=32 nprobe=32 nq=10k k=100" > /dev/null 2>&1
NQ=600000 timeout 600 ncu --set full --clock-control none -k regex:"flat_tc_kernel|tc_argmin_finish|kmeans_segment_sum" -s 2 -c 6 -f -o $O/prof_km python scripts/gpu_kmeans_assign.py > $O/r02_ncu_kmeans.log 2>&1
python scripts/ncu_summary.py $O/prof_km.ncu-rep > $O/r02_ncu_kmeans_summary.txt 2>&1; rm -f $O/prof_km.ncu-rep
timeout 400 compute-sanitizer --tool memcheck python -m pytest tests/test_flat_gpu.py -x -q -k "tensor_core_path_equals_exact_path or streaming_argmin" > $O/r02_sanitizer_flat_memcheck.log 2>&1
timeout 400 compute-sanitizer --tool racecheck python -m pytest tests/test_flat_gpu.py -x -q -k "streaming_argmin_equals_exact_path and 4096" > $O/r02_sanitizer_flat_racecheck.log 2>&1
./tests/adapter/_build/adapter_test > $O/r02_adapter_test.txt 2>&1
cuobjdump -sass faiss_b200/libfaiss_b200.so 2>/dev/null | grep -E "UTCHMMA|UTMALDG|LDTM|UTCBAR|SYNCS" | head -40 > $O/r02_sass_tcgen05_excerpt.txt
du -sh $O; ls -la $O | head -30
