"""Run-to-run reproducibility of fit / transform (random_state=42) inside one process and across processes: writes a digest of the
embeddings; a second invocation with the same path compares.  python scripts/check_determinism.py digest.npz"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tda_multimodal_b200 import umap_, workloads

path = sys.argv[1]
X = torch.from_numpy(workloads.c3_layers(n_layers=32, n=400, d=256, layers=[0, 5, 31])).cuda()   # layer 31: many components
out = {}
for rep in range(2):
    Y, st = umap_.umap_fit_batch(X[2:3], n_neighbors=15, n_components=3, metric="cosine", random_state=42, return_state=True)
    out[f"fit{rep}"] = Y.cpu().numpy()
    out[f"init{rep}"] = st["init"].cpu().numpy()
    out[f"knn{rep}"] = st["knn_indices"].cpu().numpy()
    out[f"sig{rep}"] = st["sigmas"].cpu().numpy()
    out[f"w{rep}"] = st["weight"].cpu().numpy()
    a, b = umap_.find_ab_params(1.0, 0.1)
    for l in range(2):
        T = umap_.umap_transform_batch(X[l:l + 1], X[2:3], Y, 15, metric="cosine", a=a, b=b, seed=42)
        out[f"tr{rep}_{l}"] = T.cpu().numpy()
for k in sorted(out):
    if k.endswith("1") or "1_" in k:
        k0 = k.replace("1", "0", 1) if k[:-1].endswith(("fit", "init", "knn", "sig", "w")) else k.replace("tr1", "tr0")
        print(f"in-process {k0} == {k}: {np.array_equal(out[k0], out[k])}")
if os.path.exists(path):
    ref = np.load(path)
    for k in sorted(out):
        same = np.array_equal(out[k], ref[k])
        print(f"across processes {k}: {same}" + ("" if same else f"  max|diff| {np.abs(out[k].astype(np.float64) - ref[k]).max():.3e}"))
else:
    np.savez(path, **out)
    print("wrote", path)
