"""Generates tests/golden/c2_torus_n600_dgms.npz: the CPU oracle's ripser(X, maxdim=2) diagrams for config C2 of BASELINE.json
(noisy flat torus S1 x S1 in 4096-d, workloads.c2_torus) at n=600 -- a size whose tetrahedron key space (6.5e10) spans 15 windows
of the GPU reducer, so the far-bucket path runs with its default sizes.  Distances: float64 Gram form -> float32 (the GPU path
uses its 3xTF32 tensor-core distances, so the test compares by bottleneck distance).  ~1.5 min on one core.
    python tests/golden/make_c2_golden.py [n]
"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import rips as orips
from tda_multimodal_b200 import workloads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 600      # 600: ~1 min, 1.5 GB; 1000: 5 min, 19 GB of RAM
X = workloads.c2_torus(n=n).astype(np.float64)
sq = (X * X).sum(1)
D = np.sqrt(np.maximum(sq[:, None] + sq[None] - 2.0 * X @ X.T, 0.0))
np.fill_diagonal(D, 0.0)
t = time.time()
r = orips.rips_dm(D.astype(np.float32), maxdim=2)
print(f"oracle n={n}: {time.time() - t:.1f} s; rows", [len(d) for d in r["dgms"]])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"c2_torus_n{n}_dgms.npz"), h0=r["dgms"][0], h1=r["dgms"][1], h2=r["dgms"][2],
                    diameter=np.float64(D.max()))
