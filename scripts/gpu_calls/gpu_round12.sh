#!/bin/bash
# async result copies (RipsJob) + e2e group tuning after the SGD speed-up
set -u
mkdir -p gpurun_out
python -m pytest tests/test_rips_gpu.py tests/test_dropin_gpu.py tests/test_c5_path_gpu.py -m gpu -q -x 2>&1 | tail -2 | cut -c1-200
TUNE_STEPS=10 python scripts/tune_step.py chunks=3,tail_rips_cluster=8 2>&1 | grep "min " | tee gpurun_out/tune18.log
TUNE_E2E=1 TUNE_STEPS=10 python scripts/tune_step.py chunks=4,tail_rips_cluster=8 chunks=4,tail_rips_cluster=8,tail_sgd_cluster=8 split=10-10-8-4,tail_rips_cluster=8,tail_sgd_cluster=8 split=9-9-8-6,tail_rips_cluster=8 split=8-8-7-5-4,tail_rips_cluster=8,tail_sgd_cluster=8 chunks=5,tail_rips_cluster=8 chunks=3,tail_rips_cluster=8 split=6-9-9-8,tail_rips_cluster=8 2>&1 | grep "min " | cut -c1-120 | tee gpurun_out/tune_e2e4.log
