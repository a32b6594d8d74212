#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_umap_parity_gpu.py tests/test_umap_gpu.py tests/test_rips_h2_gpu.py tests/test_dropin_gpu.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/pytest_new.log; tail -30 gpurun_out/pytest_new.log | cut -c1-300
timeout 300 python scripts/timeline.py 2>&1 | tee gpurun_out/timeline.log | head -50
timeout 900 python scripts/tune_step.py 2>&1 | tee gpurun_out/tune.log
