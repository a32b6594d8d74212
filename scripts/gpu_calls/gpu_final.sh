#!/bin/bash
# last check of the round: the driver's three GPU steps (tests, smoke, bench) on the final build
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | cut -c1-300 | tee gpurun_out/r02o_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02o_smoke.log
python bench.py > gpurun_out/r02o_bench.json 2> gpurun_out/r02o_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02o_bench.json").read().strip().splitlines()[-1])
print("layers/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 2), "launches", d["gpu_launches"], d["ms_each_step_rank0"], d["clocks"])
print("roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 4), d["roofline"]["traffic"], "cpu", d["cpu_baseline"]["value"])
PY
python scripts/time_c4_stages.py 2>&1 | grep subsets= | tee gpurun_out/r02o_c4_one_batch.log
