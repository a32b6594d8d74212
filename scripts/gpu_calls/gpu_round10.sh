#!/bin/bash
# A/B of the SGD kernel with flush-to-zero MUFU wrappers: embeddings of the old and the new library must be the same bits
set -u
mkdir -p gpurun_out
P=tda_multimodal_b200
cp $P/libtda_b200.so /tmp/new.so
cp $P/libtda_b200_old.so.bin $P/libtda_b200.so
python scripts/check_determinism.py /tmp/digest.npz 2>&1 | tail -3
TUNE_STEPS=8 python scripts/tune_step.py chunks=3,tail_rips_cluster=8 2>&1 | tail -1 | tee gpurun_out/tune15_old.log
cp /tmp/new.so $P/libtda_b200.so
python scripts/check_determinism.py /tmp/digest.npz 2>&1 | grep -v "True" | tee gpurun_out/r02m_old_vs_new_digest.log | tail -12
TUNE_STEPS=8 python scripts/tune_step.py chunks=3,tail_rips_cluster=8 chunks=3,tail_rips_cluster=8,sgd_tile=8 chunks=3,tail_rips_cluster=8,sgd_tile=32 2>&1 | tail -3 | tee gpurun_out/tune15_new.log
python -m pytest tests/test_umap_gpu.py tests/test_umap_parity_gpu.py -m gpu -q -x 2>&1 | tail -2 | cut -c1-200
