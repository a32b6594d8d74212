"""C2 (flat torus in 4096-d) at n points: wall time of the H0/H1 stage and of the H2 stage separately (+ TDA_H2_STATS=1 counters)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tda_multimodal_b200 import rips, workloads
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
X = workloads.c2_torus(n=n)
import warnings; warnings.simplefilter("ignore")
for md in (1, 2):
    torch.cuda.synchronize(); t = time.perf_counter()
    r = rips.ripser(X, maxdim=md)
    torch.cuda.synchronize(); print(f"n={n} maxdim={md}: {time.perf_counter() - t:.2f} s", [len(d) for d in r["dgms"]], flush=True)
