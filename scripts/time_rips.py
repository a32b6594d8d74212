"""Ad-hoc timing of the Rips pipeline (not the bench contract): python scripts/time_rips.py [case ...]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tda_multimodal_b200 import rips
from tests.helpers import torus3d, blobs3d

GEN = {"blobs": blobs3d, "torus": torus3d}

def run(name, n, B, reps=2):
    rng = np.random.default_rng(1)
    X = np.stack([GEN[name](n, rng) for _ in range(B)])
    pts = torch.from_numpy(X).cuda()
    dm = rips.pdist_lowdim(pts)
    torch.cuda.synchronize()
    for r in range(reps):
        t = time.perf_counter()
        res = rips.rips_batch(dm, maxdim=1, want_stats=(r == reps - 1))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
    st = res[0]["stats"]
    print(f"{name} n={n} B={B}: {dt*1e3:.1f} ms total, {dt*1e3/B:.2f} ms/cloud  stats0={st}", flush=True)

cases = sys.argv[1:] or ["blobs:1000:32", "torus:1000:1"]
for c in cases:
    name, n, B = c.split(":")
    run(name, int(n), int(B))
