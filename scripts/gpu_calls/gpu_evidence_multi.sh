#!/bin/bash
# Multi-GPU evidence (run with gpurun --gpus N): the north-star target (C3 + 256 resamples per layer, 32 layers over the ranks),
# strong scaling of config C3, config C5 at its size, fit-once / transform-many under NCCL.   usage: gpu_evidence_multi.sh N
set -u
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
( timeout 300 python scripts/run_fit_once_multirank.py gpurun_out/fit_once_w1.npz ) 2>&1 | grep "fit-once" | tee gpurun_out/r02_fit_once_n1.log
( timeout 300 $TR scripts/run_fit_once_multirank.py gpurun_out/fit_once_w1.npz ) 2>&1 | grep "fit-once" | tee gpurun_out/r02_fit_once_n$N.log
( timeout 600 $TR bench.py --gpus $N --scaling strong --workload c3c4 --steps 2 --warmup 1 --no-peaks ) > gpurun_out/r02_c3c4_strong_n$N.json 2> gpurun_out/r02_c3c4_strong_n$N.err; tail -c 800 gpurun_out/r02_c3c4_strong_n$N.json; tail -3 gpurun_out/r02_c3c4_strong_n$N.err
( timeout 600 $TR bench.py --gpus $N --scaling strong --steps 6 --warmup 3 --no-peaks ) > gpurun_out/r02_c3_strong_n$N.json 2> gpurun_out/r02_c3_strong_n$N.err; tail -c 600 gpurun_out/r02_c3_strong_n$N.json

( timeout 900 $TR scripts/run_c5.py 100000 10000 ) 2>&1 | grep "C5" | tee gpurun_out/r02_c5_100k_n$N.log
rm -f gpurun_out/fit_once_w1.npz
