#!/bin/bash
# One-GPU evidence for the configs of BASELINE.json other than the bench line: C2 (maxdim=2, n=2000), C4 (bootstrap), C5 (100k, 1 rank),
# the north-star workload c3c4 on one GPU.  Logs go to gpurun_out/ (copied into profiles/ afterwards).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_c5_path_gpu.py tests/test_rips_gpu.py -m gpu -q -x 2>&1 | tail -8 | cut -c1-300 | tee gpurun_out/r02_pytest_c4c5.log
( timeout 600 python scripts/run_c2.py 2000 ) 2>&1 | tail -5 | tee gpurun_out/r02_c2_n2000.log
( timeout 900 python bench.py --workload c3c4 --steps 2 --warmup 1 --no-cpu-baseline --no-peaks ) > gpurun_out/r02_c3c4_n1.json 2> gpurun_out/r02_c3c4_n1.err; tail -c 1500 gpurun_out/r02_c3c4_n1.json; tail -3 gpurun_out/r02_c3c4_n1.err
( timeout 900 python scripts/run_c5.py 100000 10000 ) 2>&1 | tail -8 | tee gpurun_out/r02_c5_100k_n1.log
