#!/bin/bash
# ncu evidence of the final kernels of round 2 (one GPU): launch list of the bench command + full capture of the kernels that
# changed in the last third of the round (SGD, SGD adjacency, one-launch H0, Lanczos) + the C4 subset kernels.
set -u
mkdir -p gpurun_out
TAG=r02n
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-peaks"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launch.log 2>&1
echo "launch list rc=$?"
KERNELS='sgd_cluster_kernel|sgd_adj_kernel|boruvka_chunked_kernel|lanczos_cluster_kernel|rips_sweep2_kernel'
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$KERNELS" -s 10 -c 10 -o /tmp/${TAG}_full $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full rc=$?"
REP=/tmp/${TAG}_full.ncu-rep
ncu -i $REP --page raw --csv --print-units base > gpurun_out/${TAG}_full_raw.csv 2> gpurun_out/${TAG}_export.err
ncu -i $REP --page details --csv --print-units base > gpurun_out/${TAG}_full_details.csv 2>> gpurun_out/${TAG}_export.err
ncu -i $REP --page source --csv -k regex:sgd_cluster_kernel -c 1 > gpurun_out/${TAG}_source_sgd_cluster_kernel.csv 2>> gpurun_out/${TAG}_export.err
gzip -f gpurun_out/${TAG}_source_sgd_cluster_kernel.csv
timeout 600 ncu --set full --clock-control none -k regex:'subset_|boruvka_chunked_kernel|sorted_edges' -c 8 -o /tmp/${TAG}_c4 python scripts/time_c4_stages.py > gpurun_out/${TAG}_ncu_c4.log 2>&1
ncu -i /tmp/${TAG}_c4.ncu-rep --page raw --csv --print-units base > gpurun_out/${TAG}_c4_raw.csv 2>> gpurun_out/${TAG}_export.err
ls -la gpurun_out/${TAG}_* | head -20
