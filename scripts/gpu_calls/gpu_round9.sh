#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_rips_gpu.py tests/test_rips_reducers_gpu.py tests/test_c5_path_gpu.py -m gpu -q -x 2>&1 | tail -2 | cut -c1-200
python scripts/sweep2_phase_cycles.py 2>&1 | grep "sweep2 cloud" | sort -t' ' -k3 -n | cut -c1-330 | tee gpurun_out/sweep2_phase.log | tail -4
TUNE_STEPS=8 python scripts/tune_step.py chunks=3,tail_rips_cluster=8 2>&1 | tee gpurun_out/tune14.log | tail -1
