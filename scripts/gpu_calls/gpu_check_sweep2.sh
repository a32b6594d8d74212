#!/bin/bash
# sweep2 reducer: parity tests, per-cloud statistics, bench A/B against the old reducers.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rips_reducers_gpu.py -x -q 2>&1 | tail -5 | tee gpurun_out/sweep2_tests.log
timeout 300 python scripts/rips_stats_c3.py 32 2>&1 | tee gpurun_out/sweep2_stats.log
TDA_RIPS_W0=2048 TDA_RIPS_WSPARSE=16384 timeout 300 python scripts/rips_stats_c3.py 32 2>&1 | tail -4 | tee gpurun_out/sweep2_stats_w2048.log
for mode in sweep2; do
  TDA_RIPS_REDUCER=$mode timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s2_$mode.json 2> gpurun_out/bench_s2_$mode.err
  python - "$mode" <<'PY'
import json, sys
mode = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_s2_{mode}.json").read().strip().splitlines()[-1])
    print(mode, "layers/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 1), d["roofline"]["stages_ms_per_step"])
except Exception as ex:
    print(mode, "bench failed:", ex); print(open(f"gpurun_out/bench_s2_{mode}.err").read()[-3000:])
PY
done
