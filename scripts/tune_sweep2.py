"""sweep2 option sweep on the 32 C3 clouds: time of the residual reduction (stage timer) per setting."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tda_multimodal_b200 import _lib, rips, umap_, workloads
L = _lib.lib()
X = torch.from_numpy(workloads.c3_layers(n_layers=32)).cuda()
Y = umap_.umap_fit_batch(X, n_neighbors=15, n_components=3, metric="cosine", random_state=42)
del X
dm = rips.pdist_lowdim(Y)
base = {k: _lib.get_option(k) for k in ("rips_w0", "rips_wsparse", "rips_wmax", "rips_dense_min", "rips_dense_div", "rips_cluster", "rips_warp_engine")}
def run(tag, **opts):
    for k, v in {**base, **opts}.items():
        _lib.set_option(k, v)
    ts = []
    for rep in range(3):
        L.tda_stage_timing_enable(1); L.tda_stage_timing_reset()
        rips.rips_batch(dm, maxdim=1)
        torch.cuda.synchronize()
        ts.append(_lib.stage_times()["rips_reduce"][0])
        L.tda_stage_timing_enable(0)
    print(f"{tag:50s} reduce {min(ts):7.2f} ms  (runs {[round(t, 1) for t in ts]})", flush=True)
run("default")
for w0 in (256, 512, 2048, 4096):
    run(f"w0={w0}", rips_w0=w0)
for c in (1, 2, 8):
    run(f"cluster={c}", rips_cluster=c)
for wm in (8192, 16384, 65472):
    run(f"wmax={wm}", rips_wmax=wm)
for dd in (4, 16, 32):
    run(f"dense_div={dd}", rips_dense_div=dd)
run("warp_engine=0", rips_warp_engine=0)
