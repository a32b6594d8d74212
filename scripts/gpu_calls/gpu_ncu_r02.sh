#!/bin/bash
# ncu evidence for round 2 (one GPU): launch list of the bench command + one full capture of the top kernels of one timed step.
# The .ncu-rep stays on the box (it is larger than what gpurun brings back); its pages are exported as CSV into gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r02}
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-peaks"
$CMD > gpurun_out/${TAG}_ncu_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/${TAG}_ncu_plain.log; exit 1; }
tail -c 400 gpurun_out/${TAG}_ncu_plain.log; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launch.log 2>&1
echo "launch list rc=$?"
KERNELS='rips_sweep2_kernel|sgd_cluster_kernel|lanczos_cluster_kernel|apparent_rows_kernel|pdist_gemm_kernel|knn_smooth_block_kernel|parents_kernel|sgd_adj_kernel|prep_kernel|rank_scatter_kernel'
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$KERNELS" -s 20 -c 20 -o /tmp/${TAG}_full $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full rc=$?"
REP=/tmp/${TAG}_full.ncu-rep
ls -la $REP
ncu -i $REP --page raw --csv --print-units base > gpurun_out/${TAG}_full_raw.csv 2> gpurun_out/${TAG}_export.err
ncu -i $REP --page details --csv --print-units base > gpurun_out/${TAG}_full_details.csv 2>> gpurun_out/${TAG}_export.err
for K in rips_sweep2_kernel sgd_cluster_kernel lanczos_cluster_kernel apparent_rows_kernel pdist_gemm_kernel; do
  ncu -i $REP --page source --csv -k regex:$K -c 1 > gpurun_out/${TAG}_source_$K.csv 2>> gpurun_out/${TAG}_export.err
  gzip -f gpurun_out/${TAG}_source_$K.csv
done
SZ=$(stat -c %s $REP)
if [ "$SZ" -lt 30000000 ]; then cp $REP gpurun_out/; fi
ls -la gpurun_out/ | head -40
du -sh gpurun_out
