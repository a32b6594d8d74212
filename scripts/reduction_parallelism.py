"""How much parallelism is there in the residual H1 reduction of a C3 cloud?  CPU oracle only (UMAP oracle -> Rips oracle with
dependency statistics): total pivot steps of the reduced columns against the heaviest chain of columns that wait for each other."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import umap_oracle as uo, rips as orips
from tda_multimodal_b200 import workloads
layers = [int(x) for x in sys.argv[1:]] or [0, 2]
for l in layers:
    X = workloads.c3_layer(l)
    Y = uo.UMAPOracle(n_neighbors=15, n_components=3, min_dist=0.1, metric="cosine", random_state=42).fit_transform(X)
    t = time.time()
    r = orips.ripser(Y, maxdim=1, with_stats=True) if "with_stats" in orips.ripser.__code__.co_varnames else None
    if r is None:
        r = orips.rips_dm(orips.euclidean_dm_f32(Y), maxdim=1, with_stats=True)
    s = r["stats"][1]
    print(f"layer {l}: rips {time.time() - t:.1f} s; reduced columns {s['reduced']}, additions {s['additions']}, pivot steps of reduced columns {s['dep_total_steps']}, "
          f"heaviest dependency chain {s['dep_critical_steps']} steps over {s['dep_depth']} columns -> ideal speed-up {s['dep_total_steps'] / max(1, s['dep_critical_steps']):.1f}x", flush=True)
