"""GPU: the reference's call sequence through the import shims (`import umap`, `from ripser import ripser`,
`from persim import plot_diagrams`), as debug_tda_pipeline.py:59-157 makes it, on the C1 workload (48-sample 6x6
set, 36 'bound' samples, 32 layers x 4096-d).  /root/reference is not available on the GPU box, so the loop body is
restated here call by call; what is checked is the boundary: dtypes, shapes, JSON-serialisability, schema."""
import json
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_debug_pipeline_call_sequence(tmp_path):
    import torch
    sys.path.insert(0, os.path.join(ROOT, "shims"))
    import umap
    from ripser import ripser
    from sklearn.metrics import silhouette_score
    from tda_multimodal_b200 import workloads
    from tests.helpers import get_persistence

    with open(os.path.join(ROOT, "tests", "golden", "ref_metadata_6x6.json")) as f:
        metadata = json.load(f)
    data = workloads.c1_activations(metadata=metadata)
    all_data = {k: {"metadata": v["metadata"], "activations": {lk: torch.from_numpy(lv) for lk, lv in v["activations"].items()}} for k, v in data.items()}
    torch.save(all_data, tmp_path / "all_activations.pt")
    all_data = torch.load(tmp_path / "all_activations.pt")
    sample_ids = sorted(item["id"] for item in metadata if item["type"] == "bound")
    assert len(sample_ids) == 36
    id_to_meta = {item["id"]: item for item in metadata}
    shape_labels = [id_to_meta[s]["shape"] for s in sample_ids]
    stats = []
    for i in range(0, 32, 4):
        cloud = torch.stack([all_data[s]["activations"][f"layer_{i}"] for s in sample_ids]).numpy().astype(np.float64)
        reducer = umap.UMAP(n_neighbors=6, n_components=3, min_dist=0.1, random_state=42, metric="cosine")
        low = reducer.fit_transform(cloud)
        assert low.shape == (36, 3) and low.dtype == np.float32
        np.save(tmp_path / f"layer_{i}_cloud.npy", low)
        dgms = ripser(low, maxdim=1)["dgms"]
        assert all(d.dtype == np.float64 and d.ndim == 2 and d.shape[1] == 2 for d in dgms)
        h0, m0 = get_persistence(dgms[0])
        h1, m1 = get_persistence(dgms[1])
        stats.append({"layer": i, "n_h1_features": len(h1), "max_h1_persistence": m1, "all_h1_persistence_values": h1.tolist(),
                      "n_h0_features": len(dgms[0]) - len(h0), "max_h0_persistence": m0,
                      "silhouette_shape": silhouette_score(low, shape_labels)})
    text = json.dumps(stats, indent=2)  # np.float64 maxima serialise; float32 would raise (SURVEY.md 8b "JSON gotcha")
    back = json.loads(text)
    assert len(back) == 8 and all(r["n_h0_features"] == 1 for r in back)


def test_persim_shim_accepts_dgms():
    sys.path.insert(0, os.path.join(ROOT, "shims"))
    from persim import bottleneck
    from ripser import ripser
    rng = np.random.default_rng(0)
    t = rng.uniform(0, 2 * np.pi, 120)
    X = np.c_[np.cos(t), np.sin(t)].astype(np.float32)
    d = ripser(X, maxdim=1)["dgms"]
    assert bottleneck(d[1], d[1]) == 0.0
    d2 = ripser(X * np.float32(1.01), maxdim=1)["dgms"]
    assert bottleneck(d[1], d2[1]) <= 0.011 * 2
