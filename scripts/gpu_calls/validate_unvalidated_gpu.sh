#!/bin/bash
# Round 1 ended with two pieces written after the GPU minutes were spent (compiled, CPU-modelled, never run on a B200).
# Run this first on a GPU box (e.g. `gpurun --timeout 600 -- 'bash scripts/validate_unvalidated_gpu.sh'`); outputs in gpurun_out/.
#   1. substitute-then-verify sweep reducer (TDA_RIPS_REDUCER=verify): parity tests, then an A/B of bench.py
#   2. pipeline.fit_once_transform_many on one rank
#   3. TDA_SGD_AGG=1 (per-epoch SGD kernel with warp-aggregated updates of the slot block's own vertex): parity test + A/B
set -u
mkdir -p gpurun_out
export TDA_TEST_UNVALIDATED=1
timeout 400 python -m pytest tests/test_zz_verify_reducer_gpu.py tests/test_zz_fit_once_gpu.py tests/test_rips_h2_gpu.py -q 2>&1 | tail -15 | tee gpurun_out/unvalidated_tests.log
# the default reducer's own parity tests, run with the new reducer selected (bit-exact diagrams + simplex pairs)
TDA_RIPS_REDUCER=verify timeout 300 python -m pytest tests/test_rips_gpu.py -q -x 2>&1 | tail -5 | tee gpurun_out/verify_on_default_tests.log
for mode in sweep verify; do
  TDA_RIPS_REDUCER=$mode timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab_$mode.json 2> gpurun_out/bench_ab_$mode.err
  python - "$mode" <<'PY'
import json, sys
mode = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_ab_{mode}.json").read().strip().splitlines()[-1])
    print(mode, "layers/s", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 1), "reduce ms", d["roofline"]["stages_ms_per_step"]["rips_reduce"])
except Exception as ex:
    print(mode, "bench failed:", ex)
PY
done
TDA_SGD_AGG=1 timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab_sgdagg.json 2> gpurun_out/bench_ab_sgdagg.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_ab_sgdagg.json").read().strip().splitlines()[-1])
    print("sgd_agg layers/s", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 1), "sgd ms", d["roofline"]["stages_ms_per_step"]["umap_sgd"])
except Exception as ex:
    print("sgd_agg bench failed:", ex)
PY
