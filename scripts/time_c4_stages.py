"""Config C4, ONE batch alone (256 resamples of 1000 points of one 2000-point cloud): stage times without a second batch in
flight, with and without the subset front end.  python scripts/time_c4_stages.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tda_multimodal_b200 import pipeline, workloads, _lib
L = _lib.lib()
Y = pipeline.layer_sweep(torch.from_numpy(workloads.c3_layers(n_layers=32, layers=[0])).cuda())["embedding"].contiguous()
for sub in (True, False):
    for rep in range(3):
        L.tda_stage_timing_reset(); L.tda_stage_timing_enable(1)
        torch.cuda.synchronize(); t = time.perf_counter()
        res = pipeline.bootstrap_rips(Y, n_resamples=256, size=1000, max_batch=256, subsets=sub)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
        st = {k: round(v[0], 2) for k, v in _lib.stage_times().items() if v[0] > 0.01}
        L.tda_stage_timing_enable(0)
    print(f"subsets={sub}: {dt*1e3:.1f} ms wall for one batch; device stages (ms): {st}; sum {sum(st.values()):.1f}")
