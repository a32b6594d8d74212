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

# With the oracle's apparent-pair shortcut (Ripser 1.2's; rips_oracle.cpp header): 600: 5 s, 0.25 GB; 1000: 33 s, 2.3 GB; 2000 (the
# full size of config C2): see the printed time.  Without it (how the n=600 / n=1000 files were first made; the shortcut reproduces
# both files bit for bit): 600: ~1 min, 1.5 GB; 1000: 5 min, 19 GB; 2000: out of memory.
n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
X = workloads.c2_torus(n=n).astype(np.float64)
sq = (X * X).sum(1)
D = np.sqrt(np.maximum(sq[:, None] + sq[None] - 2.0 * X @ X.T, 0.0))
np.fill_diagonal(D, 0.0)
t = time.time()
r = orips.rips_dm(D.astype(np.float32), maxdim=2, apparent=True, with_stats=True)
print(f"oracle n={n}: {time.time() - t:.1f} s; rows", [len(d) for d in r["dgms"]], "num_edges", r["num_edges"], "stats", r["stats"][1:])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"c2_torus_n{n}_dgms.npz"), h0=r["dgms"][0], h1=r["dgms"][1], h2=r["dgms"][2],
                    diameter=np.float64(D.max()), num_edges=np.int64(r["num_edges"]))
