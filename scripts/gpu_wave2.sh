#!/bin/bash
# round-2 wave 2: K-split kernel / large k / fp16 storage / paging / bfKnn -- tests first, then timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_flat_gpu.py -x -q -m gpu --durations=5 > gpurun_out/r02_wave2_flat.log 2>&1; echo "flat rc=$?"
tail -15 gpurun_out/r02_wave2_flat.log
timeout 600 python -m pytest tests/test_ivf_gpu.py -q -m gpu -k "k2048 or shards_ivf" > gpurun_out/r02_wave2_ivf.log 2>&1; echo "ivf rc=$?"
tail -5 gpurun_out/r02_wave2_ivf.log
timeout 600 python scripts/gpu_widepath.py > gpurun_out/r02_widepath.log 2>&1; echo "wide rc=$?"
cat gpurun_out/r02_widepath.log | tail -12
