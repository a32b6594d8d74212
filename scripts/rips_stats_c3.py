"""Per-cloud reducer statistics on the C3 workload (which clouds are slow, and in which phase).  Works with every reducer
(TDA_RIPS_REDUCER=sweep2|sweep|verify|bitset); cycle counters are SM cycles of thread 0 of the cloud's CTA."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tda_multimodal_b200 import workloads, umap_, rips, _lib
L = int(sys.argv[1]) if len(sys.argv) > 1 else 32
X = torch.from_numpy(workloads.c3_layers(n_layers=32, layers=range(L))).cuda()
Y = umap_.umap_fit_batch(X, n_neighbors=15, n_components=3, metric="cosine", random_state=42)
del X
dm = rips.pdist_lowdim(Y)
for rep in range(3):
    torch.cuda.synchronize(); t = time.perf_counter()
    res = rips.rips_batch(dm, maxdim=1, want_stats=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
print(f"reducer {_lib.rips_reducer()}: rips_batch {L} clouds: {dt*1e3:.1f} ms (whole Rips stage incl. sort / H0 / apparent pairs / D2H)")
keys = [k for k in res[0]["stats"] if not k.startswith("spare")]
keys = [k for k in keys if k not in ("columns", "apparent", "edges_via_columns", "max_v")]
cyc = [k for k in keys if k.startswith("cyc_")]
rows = sorted(range(L), key=lambda p: -sum(res[p]["stats"][k] for k in cyc))
print("cloud " + " ".join(f"{k[:13]:>13s}" for k in keys) + "   total_Mcyc")
for p in rows[:10] + rows[-2:]:
    st = res[p]["stats"]
    print(f"{p:5d} " + " ".join(f"{st[k]:13d}" for k in keys) + f"   {sum(st[k] for k in cyc) / 1e6:8.2f}")
tot = {k: sum(r["stats"][k] for r in res) for k in keys}
print("sum   " + " ".join(f"{tot[k]:13d}" for k in keys))
