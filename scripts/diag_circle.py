import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import umap_oracle as uo, rips as orips
from tda_multimodal_b200.umap_ import UMAP
from tda_multimodal_b200 import workloads
from sklearn.manifold import trustworthiness
for seed in (77, 78):
    rng = np.random.default_rng(seed)
    t = rng.uniform(0, 2 * np.pi, 400)
    X = workloads._embed(np.c_[np.cos(t), np.sin(t)], 512, rng, noise=0.002, scale=5.0)
    um = UMAP(n_neighbors=15, n_components=3, random_state=42, metric="cosine")
    Yg = um.fit_transform(X)
    Y0 = um._state["init"][0].cpu().numpy()
    uo_ = uo.UMAPOracle(n_neighbors=15, n_components=3, metric="cosine", random_state=42)
    Yo = uo_.fit_transform(X)
    for name, Y in (("gpu", Yg), ("oracle", Yo), ("gpu_init", Y0), ("oracle_init", uo_._init_embedding)):
        d1 = orips.ripser(Y, maxdim=1)["dgms"][1]
        pers = np.sort(d1[:, 1] - d1[:, 0])[::-1]
        print(seed, name, "top pers", np.round(pers[:3], 3), "trust", round(trustworthiness(X, Y, n_neighbors=10, metric="cosine"), 4), "range", np.round(np.ptp(Y, 0), 2), flush=True)
