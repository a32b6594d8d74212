"""GPU: analyze_tda_over_layers.py's calling convention (fit on one layer, transform the others, Rips on every embedding)
through pipeline.fit_once_transform_many on one rank (the state broadcast is covered on the CPU by tests/test_multirank_cpu.py
and under NCCL by scripts/run_fit_once_multirank.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_fit_once_transform_many_single_rank():
    import torch
    from sklearn.manifold import trustworthiness
    from oracle import rips as orips
    from tda_multimodal_b200 import pipeline, workloads
    X = torch.from_numpy(workloads.c3_layers(layers=[0, 10, 31], n=400, d=256)).cuda()
    emb, dgms = pipeline.fit_once_transform_many(X, fit_layer=0, n_neighbors=15)
    emb_last, _ = pipeline.fit_once_transform_many(X[:2], n_neighbors=15)   # default: fit on the last layer (analyze_tda_over_layers.py:68)
    assert sorted(emb_last) == [0, 1]
    assert sorted(emb) == [0, 1, 2] and len(dgms) == 3
    for l in range(3):
        Y = emb[l].cpu().numpy()
        assert Y.shape == (400, 3) and np.isfinite(Y).all()
        want = orips.ripser(Y, maxdim=1)["dgms"]
        assert np.array_equal(dgms[l][0], want[0]) and np.array_equal(dgms[l][1], want[1])
    assert trustworthiness(X[0].cpu().numpy(), emb[0].cpu().numpy(), n_neighbors=10, metric="cosine") > 0.85
