#!/bin/bash
# One GPU round trip: the -m gpu suite (or the tests named in $TESTS), the per-cloud reducer statistics, a short bench and
# (NCU=1) an ncu capture of the reducer.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest ${TESTS:-tests} -m gpu -q ${PYTEST_ARGS:--x} 2>&1 | tail -${TAIL:-15} | tee gpurun_out/pytest_gpu.log
timeout 300 python scripts/rips_stats_c3.py 32 2>&1 | tee gpurun_out/sweep2_stats.log | cut -c1-330
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
    print("layers/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 1), "launches", d["gpu_launches"], d["roofline"]["stages_ms_per_step"])
except Exception as ex:
    print("bench failed:", ex); print(open("gpurun_out/bench_quick.err").read()[-3000:])
PY
if [ "${NCU:-0}" = "1" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:rips_sweep2 -c 1 -o gpurun_out/ncu_sweep2 -f python scripts/rips_stats_c3.py 32 > gpurun_out/ncu_sweep2.log 2>&1
  tail -3 gpurun_out/ncu_sweep2.log
fi
