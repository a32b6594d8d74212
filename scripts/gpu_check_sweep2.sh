#!/bin/bash
# sweep2 reducer: parity tests, then bench A/B against the old reducers.  gpurun --timeout 1200 -- 'bash scripts/gpu_check_sweep2.sh'
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rips_reducers_gpu.py -x -q 2>&1 | tail -15 | tee gpurun_out/sweep2_tests.log
timeout 600 python -m pytest tests/test_rips_gpu.py tests/test_rips_h2_gpu.py -x -q 2>&1 | tail -8 | tee gpurun_out/sweep2_rips_tests.log
for mode in sweep2 verify; do
  TDA_RIPS_REDUCER=$mode timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s2_$mode.json 2> gpurun_out/bench_s2_$mode.err
  tail -c 1500 gpurun_out/bench_s2_$mode.json
  echo
done
