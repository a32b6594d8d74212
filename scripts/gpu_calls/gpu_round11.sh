#!/bin/bash
# A/B of sgd_adj_kernel on 4 CTAs per cloud + unrolled slot-table passes (SGD adjacency, Lanczos set-up): same bits as the old library
set -u
mkdir -p gpurun_out
P=tda_multimodal_b200
cp $P/libtda_b200.so /tmp/new.so
cp $P/libtda_b200_old.so.bin $P/libtda_b200.so
python scripts/check_determinism.py /tmp/digest.npz 2>&1 | tail -1
cp /tmp/new.so $P/libtda_b200.so
python scripts/check_determinism.py /tmp/digest.npz 2>&1 | grep -v "True" | tee gpurun_out/r02m_old_vs_new_digest.log | tail -12
echo "digest lines not True: $(wc -l < gpurun_out/r02m_old_vs_new_digest.log)"
TUNE_STEPS=8 python scripts/tune_step.py chunks=3,tail_rips_cluster=8 chunks=3,tail_rips_cluster=8 2>&1 | grep "min " | tee gpurun_out/tune17.log
python scripts/lanczos_phase_cycles.py 2>&1 | grep "lanczos cloud" | head -8 | cut -c1-400 | tee gpurun_out/r02m_lanczos_phase.log
