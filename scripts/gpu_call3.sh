#!/bin/bash
set -x
mkdir -p gpurun_out
./scripts/ubench/ubench_pq -1 8 24416 > gpurun_out/r02_ubench_pq.txt 2>&1
./scripts/ubench/ubench_pq -1 4 1920 > gpurun_out/r02_ubench_pq_short.txt 2>&1
cat gpurun_out/r02_ubench_pq.txt gpurun_out/r02_ubench_pq_short.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ivfpq_scan_interleaved -s 3 -c 1 -o gpurun_out/r02_prof_ivfpq_v2 -f python bench_ivf.py --index ivfpq --steps 1 --warmup 3 --recall-queries 0 > gpurun_out/r02_ncu_ivfpq_v2.log 2>&1
tail -3 gpurun_out/r02_ncu_ivfpq_v2.log
ls -la gpurun_out/*.ncu-rep | tail -3
