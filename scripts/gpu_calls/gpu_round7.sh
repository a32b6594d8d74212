#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | cut -c1-300 | tee gpurun_out/r02f_pytest_gpu.log
python scripts/check_determinism.py /tmp/det.npz > /dev/null 2>&1; python scripts/check_determinism.py /tmp/det.npz 2>&1 | grep -c True; python scripts/check_determinism.py /tmp/det.npz 2>&1 | grep False
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02f_bench.json").read().strip().splitlines()[-1])
print("layers/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 2), "launches", d["gpu_launches"], d["ms_each_step_rank0"])
for k, v in d["roofline"]["stages"].items():
    print(f"  {k:16s} sum {v['sum_ms_per_step']:8.3f} wall {v['wall_ms_per_step']:8.3f} launches {v['launches_per_step']:6.1f} achieved {v['achieved']:10.2f} {v['unit']:8s} frac {v['frac']:.4f}")
print(d["roofline"]["secondary"])
PY
