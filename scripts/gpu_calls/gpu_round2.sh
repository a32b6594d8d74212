#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rips_reducers_gpu.py tests/test_umap_gpu.py -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
timeout 300 python scripts/rips_stats_c3.py 32 2>&1 | tee gpurun_out/sweep2_stats.log | cut -c1-330
timeout 300 python scripts/timeline.py 2>&1 | tee gpurun_out/timeline.log
timeout 600 python scripts/tune_sweep2.py 2>&1 | tee gpurun_out/tune_sweep2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sgd_cluster -c 1 -o gpurun_out/ncu_sgd -f python scripts/timeline.py > gpurun_out/ncu_sgd.log 2>&1
tail -2 gpurun_out/ncu_sgd.log
