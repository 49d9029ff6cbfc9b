#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 300 python -m pytest tests/test_flat_gpu.py -x -q -m gpu > gpurun_out/r02_wave3_flat.log 2>&1; echo "flat rc=$?"; tail -3 gpurun_out/r02_wave3_flat.log
timeout 600 python scripts/gpu_widepath.py > gpurun_out/r02_widepath.log 2>&1; echo "wide rc=$?"; tail -9 gpurun_out/r02_widepath.log
NQ=2000000 timeout 300 python scripts/gpu_kmeans_assign.py > gpurun_out/r02_kmeans_assign2.log 2>&1; echo "kmeans rc=$?"; head -4 gpurun_out/r02_kmeans_assign2.log
timeout 900 python scripts/gpu_sweep_flat.py > gpurun_out/r02_sweep_flat.log 2>&1; echo "sweep rc=$?"; cat gpurun_out/r02_sweep_flat.log
