#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | cut -c1-300 | tee gpurun_out/r02i_pytest_gpu.log
for cfg in "" "TDA_SGD_CLUSTER=4 TDA_RIPS_CLUSTER=4"; do
  echo "== 4 layers, $cfg"
  env $cfg python bench.py --layers 4 --steps 6 --warmup 3 --no-cpu-baseline --no-peaks 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), d['ms_each_step_rank0'], {k:round(v['sum_ms_per_step'],2) for k,v in d['roofline']['stages'].items()})"
done
python bench.py --steps 8 --warmup 3 > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02i_bench.json").read().strip().splitlines()[-1])
print("layers/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 1), "launches", d["gpu_launches"], d["ms_each_step_rank0"], d["clocks"])
for k, v in d["roofline"]["stages"].items():
    print(f"  {k:16s} sum {v['sum_ms_per_step']:8.3f} wall {v['wall_ms_per_step']:8.3f} launches {v['launches_per_step']:6.1f} achieved {v['achieved']:10.2f} {v['unit']:8s} frac {v['frac']:.4f} traffic {v.get('traffic')}")
print("  cpu", d.get("cpu_baseline"))
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02i_bench_reference.json 2> gpurun_out/r02i_bench_reference.err; tail -c 700 gpurun_out/r02i_bench_reference.json
