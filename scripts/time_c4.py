"""Config C4 timing: bootstrap resamples of a 3-D cloud, Rips H0/H1, batched (resamples/s on one GPU)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tda_multimodal_b200 import pipeline
FIXED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "c3_Y_fixed.npy")
if os.path.exists(FIXED):
    Y = torch.from_numpy(np.load(FIXED)[:2]).cuda()
else:   # the UMAP embeddings of the first two C3 layers (what the north-star workload resamples)
    from tda_multimodal_b200 import workloads
    Y = pipeline.layer_sweep(torch.from_numpy(workloads.c3_layers(n_layers=32, layers=[0, 1])).cuda())["embedding"].contiguous()
R = int(sys.argv[1]) if len(sys.argv) > 1 else 256
SUB = None if os.environ.get("TDA_C4_SUBSETS", "") == "" else os.environ["TDA_C4_SUBSETS"] != "0"   # A/B of the subset front end
from tda_multimodal_b200 import _lib
L = _lib.lib()
for rep in range(3):
    if rep == 2:
        L.tda_stage_timing_reset(); L.tda_stage_timing_enable(1)
    torch.cuda.synchronize(); t = time.perf_counter()
    res = pipeline.bootstrap_rips(Y, n_resamples=R, size=1000, max_batch=256, subsets=SUB)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
print("stage sums (ms):", {k: round(v[0], 1) for k, v in _lib.stage_times().items() if v[0] > 0.05})
n1 = np.mean([len(r["dgms"][1]) for r in res[0]])
print(f"C4: 2 layers x {R} resamples x 1000 pts: {dt*1e3:.1f} ms  = {2*R/dt:.1f} resamples/s  (mean H1 rows {n1:.1f})")
