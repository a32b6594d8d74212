"""Per-cloud reducer statistics on the C3 workload (which clouds are slow, and in which phase)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tda_multimodal_b200 import workloads, umap_, rips
L = int(sys.argv[1]) if len(sys.argv) > 1 else 32
FIXED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "c3_Y_fixed.npy")
if os.path.exists(FIXED) and not os.environ.get("FRESH_UMAP"):
    Y = torch.from_numpy(np.load(FIXED)[:L]).cuda()   # a saved GPU UMAP output: identical reducer work from run to run
else:
    X = torch.from_numpy(workloads.c3_layers(n_layers=32, layers=range(L))).cuda()
    Y = umap_.umap_fit_batch(X, n_neighbors=15, n_components=3, metric="cosine", random_state=42)
dm = rips.pdist_lowdim(Y)
for rep in range(2):
    torch.cuda.synchronize(); t = time.perf_counter()
    res = rips.rips_batch(dm, maxdim=1, want_stats=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
print(f"rips_batch {L} clouds: {dt*1e3:.1f} ms")
keys = ["reduced", "additions", "pushes", "pops", "extensions", "max_v", "cyc_extract", "cyc_owner", "cyc_gen", "cyc_badd", "cyc_ext", "cyc_final", "ext_edges", "badd_edges"]
rows = sorted(range(L), key=lambda p: -(res[p]["stats"]["cyc_extract"] + res[p]["stats"]["cyc_gen"] + res[p]["stats"]["cyc_ext"] + res[p]["stats"]["cyc_final"] + res[p]["stats"]["cyc_badd"]))
print("cloud " + " ".join(f"{k:>11s}" for k in keys))
for p in rows[:8] + rows[-2:]:
    st = res[p]["stats"]
    print(f"{p:5d} " + " ".join(f"{st[k]:11d}" for k in keys))
if os.environ.get("SAVE_Y"):
    np.save(os.environ["SAVE_Y"], Y.cpu().numpy())
