"""Config C2 of BASELINE.json: noisy flat torus S1 x S1, n points embedded in 4096-d, ripser(X, maxdim=2) on the raw distance
matrix (tensor-core pdist -> Rips H0/H1/H2).  Expect two long H1 bars and one long H2 bar."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tda_multimodal_b200 import rips, workloads
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
X = workloads.c2_torus(n=n)
for rep in range(int(os.environ.get("C2_REPS", "1"))):
    torch.cuda.synchronize(); t = time.perf_counter()
    r = rips.ripser(X, maxdim=2)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
d = r["dgms"]
p1 = np.sort(d[1][:, 1] - d[1][:, 0])[::-1]; p2 = np.sort(d[2][:, 1] - d[2][:, 0])[::-1]
print(f"C2 n={n}: {dt:.2f} s; num_edges {r['num_edges']}; H1 rows {len(d[1])} top {np.round(p1[:4],3)}; H2 rows {len(d[2])} top {np.round(p2[:3],3)}")
