#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_rips_gpu.py tests/test_rips_reducers_gpu.py -m gpu -q -x 2>&1 | tail -2 | cut -c1-200
TUNE_STEPS=6 python scripts/tune_step.py chunks=3,tail_rips_cluster=8 chunks=3,tail_rips_cluster=8,rips_wc_max_rows=65536 chunks=3,tail_rips_cluster=8,rips_wc_max_rows=32768 chunks=3,tail_rips_cluster=8,rips_wc_max_rows=16384 chunks=3,tail_rips_cluster=8,rips_wc_max_rows=131072 2>&1 | tee gpurun_out/tune13.log | tail -5
for i in 1 2; do
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-peaks 2>/dev/null > gpurun_out/bench_try$i.json
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_try$i.json").read().strip().splitlines()[-1])
print(round(d["value"], 1), round(d["e2e"]["value"], 1), d["ms_each_step_rank0"], d["clocks"]["samples"])
PY
done
