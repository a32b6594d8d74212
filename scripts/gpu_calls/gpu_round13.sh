#!/bin/bash
# subset front end (C4) + H0 from the sorted edge list: parity tests, C4 timing A/B, C3 step
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_rips_subsets_gpu.py tests/test_rips_gpu.py tests/test_rips_reducers_gpu.py tests/test_c5_path_gpu.py tests/test_dropin_gpu.py tests/test_rips_h2_gpu.py -m gpu -q -x 2>&1 | tail -12 | cut -c1-300
python scripts/time_c4.py 2>&1 | tail -2 | tee gpurun_out/r02n_c4_subsets.log
TDA_C4_SUBSETS=0 python scripts/time_c4.py 2>&1 | tail -2 | tee gpurun_out/r02n_c4_per_resample.log
TUNE_STEPS=8 python scripts/tune_step.py chunks=3,tail_rips_cluster=8 chunks=3,tail_rips_cluster=8,rips_h0_chunked=0 2>&1 | grep "min " | tee gpurun_out/tune19.log
