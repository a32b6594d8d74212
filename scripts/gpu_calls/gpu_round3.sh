#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
timeout 300 python scripts/timeline.py 2>&1 | tee gpurun_out/timeline.log
TDA_SWEEP_CHUNKS=4 timeout 300 python scripts/timeline.py 2>&1 | head -3 | tee gpurun_out/timeline4.log
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_full.json").read().strip().splitlines()[-1])
    print("layers/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 1), "launches", d["gpu_launches"])
    for k, v in d["roofline"]["stages"].items():
        print(f"  {k:16s} sum {v['sum_ms_per_step']:8.3f} wall {v['wall_ms_per_step']:8.3f} launches {v['launches_per_step']:6.1f} achieved {v['achieved']:10.2f} {v['unit']:8s} frac {v['frac']:.4f}")
    print("  peaks", d["roofline"]["peaks"])
    print("  cpu", d.get("cpu_baseline"))
except Exception as ex:
    print("bench failed:", ex); print(open("gpurun_out/bench_full.err").read()[-3000:])
PY
