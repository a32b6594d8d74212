"""Timing of ripser(X, maxdim=2) on the GPU vs the CPU oracle for growing clouds."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tda_multimodal_b200 import rips
from tests.helpers import torus3d
ns = [int(a) for a in sys.argv[1:]] or [100, 200, 400, 800]
for n in ns:
    X = torus3d(n, np.random.default_rng(1))
    for rep in range(2):
        torch.cuda.synchronize(); t = time.perf_counter()
        d = rips.ripser(X, maxdim=2)["dgms"]
        torch.cuda.synchronize(); dt = time.perf_counter() - t
    msg = f"n={n}: GPU {dt*1e3:.1f} ms  H1 rows {len(d[1])}  H2 rows {len(d[2])}  top H2 pers {np.sort(d[2][:,1]-d[2][:,0])[::-1][:2] if len(d[2]) else []}"
    if n <= 200 and os.environ.get("ORACLE", "1") == "1":
        from oracle import rips as orips
        t = time.perf_counter(); w = orips.ripser(X, maxdim=2)["dgms"]; dto = time.perf_counter() - t
        msg += f"  | oracle {dto*1e3:.0f} ms, H2 rows {len(w[2])}"
    print(msg, flush=True)
