"""Config C4 timing: bootstrap resamples of a 3-D cloud, Rips H0/H1, batched (resamples/s on one GPU)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tda_multimodal_b200 import pipeline
FIXED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "c3_Y_fixed.npy")
Y = torch.from_numpy(np.load(FIXED)[:2]).cuda()
R = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for rep in range(2):
    torch.cuda.synchronize(); t = time.perf_counter()
    res = pipeline.bootstrap_rips(Y, n_resamples=R, size=1000, max_batch=256)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
n1 = np.mean([len(r["dgms"][1]) for r in res[0]])
print(f"C4: 2 layers x {R} resamples x 1000 pts: {dt*1e3:.1f} ms  = {2*R/dt:.1f} resamples/s  (mean H1 rows {n1:.1f})")
