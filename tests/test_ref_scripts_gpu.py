"""GPU: the reference's three analysis scripts, UNMODIFIED, executed with runpy against this repository's import shims
(`import umap`, `from ripser import ripser`, `from persim import plot_diagrams`) -- the drop-in contract of north_star.
The scripts are test fixtures copied by tests/golden/make_ref_scripts.py into tests/golden/_ref_scripts/ (git-ignored, travels
with the working tree); inputs are synthetic all_activations.pt files of the shapes the scripts expect; matplotlib is a stub
(tests/stubs).  Checked: the scripts run to completion, summary_stats.json has the reference's schema and key order and is
valid JSON, the saved point clouds are float32 [N,3], every layer was processed."""
import json
import os
import runpy
import shutil
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "tests", "golden", "_ref_scripts")
needs_scripts = pytest.mark.skipif(not os.path.isdir(REF), reason="tests/golden/_ref_scripts missing: run tests/golden/make_ref_scripts.py where /root/reference exists")


def _torch_dict(data):
    import torch
    return {k: {"metadata": v["metadata"], "activations": {lk: torch.from_numpy(lv) for lk, lv in v["activations"].items()}} for k, v in data.items()}


def _run(script_rel, workdir, monkeypatch):
    """Execute one reference script as __main__ inside `workdir` (its relative paths resolve there)."""
    dst = os.path.join(workdir, script_rel)
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    shutil.copyfile(os.path.join(REF, script_rel), dst)
    monkeypatch.chdir(workdir)
    for p in (os.path.join(ROOT, "shims"), os.path.join(ROOT, "tests", "stubs")):
        monkeypatch.syspath_prepend(p)
    for mod in ("umap", "ripser", "persim", "matplotlib", "matplotlib.pyplot"):
        monkeypatch.delitem(sys.modules, mod, raising=False)
    return runpy.run_path(dst, run_name="__main__")


@needs_scripts
def test_debug_tda_pipeline_script(tmp_path, monkeypatch):
    import torch
    from tda_multimodal_b200 import workloads
    with open(os.path.join(ROOT, "tests", "golden", "ref_metadata_6x6.json")) as f:
        metadata = json.load(f)
    d = tmp_path / "data" / "physics_experiment_6x6"
    d.mkdir(parents=True)
    torch.save(_torch_dict(workloads.c1_activations(metadata=metadata)), d / "all_activations.pt")
    with open(d / "metadata.json", "w") as f:
        json.dump(metadata, f)
    _run("debug_tda_pipeline.py", str(tmp_path), monkeypatch)
    out = tmp_path / "tda_debug_output"
    with open(out / "summary_stats.json") as f:
        stats = json.load(f)
    assert len(stats) == 32
    want_keys = ["layer", "n_h1_features", "max_h1_persistence", "all_h1_persistence_values", "n_h0_features", "max_h0_persistence",
                 "silhouette_shape", "silhouette_color"]
    for i, rec in enumerate(stats):
        assert list(rec) == want_keys and rec["layer"] == i
        assert rec["n_h0_features"] >= 1 and rec["n_h1_features"] == len(rec["all_h1_persistence_values"])
        assert -1.0 <= rec["silhouette_shape"] <= 1.0
    for i in (0, 17, 31):
        cloud = np.load(out / "point_clouds_3d" / f"layer_{i}_cloud.npy")
        assert cloud.shape == (36, 3) and cloud.dtype == np.float32 and np.isfinite(cloud).all()
        assert (out / "diagrams" / f"layer_{i}_diagram.png").exists()


@needs_scripts
def test_analyze_tda_over_layers_script(tmp_path, monkeypatch):
    """fit on the last layer, transform every layer (n_neighbors = 36 // 2 = 18: the general kNN kernel), ripser on each."""
    import torch
    from tda_multimodal_b200 import workloads
    with open(os.path.join(ROOT, "tests", "golden", "ref_metadata_6x6.json")) as f:
        metadata = json.load(f)
    d = tmp_path / "data" / "physics_experiment"
    d.mkdir(parents=True)
    torch.save(_torch_dict(workloads.c1_activations(metadata=metadata)), d / "all_activations.pt")
    ns = _run("analyze_tda_over_layers.py", str(tmp_path), monkeypatch)
    res = ns["results_per_layer"]
    assert len(res) == 32
    for r in res:
        assert r["dgms"][0].dtype == np.float64 and r["dgms"][0].shape[1] == 2 and r["dgms"][1].shape[1] == 2
        assert np.isinf(r["dgms"][0][-1, 1])
    assert len(ns["n_loops_per_layer"]) == 32 and all(np.isfinite(v) for v in ns["max_h0_persistence"])


@needs_scripts
def test_analyze_adversarial_tda_script(tmp_path, monkeypatch):
    """four conditions (36 / 180 / 180 / 324 samples), 32 layers each, four silhouette scores per layer."""
    import torch
    from tda_multimodal_b200 import workloads
    d = tmp_path / "data" / "physics_experiment_6x6"
    d.mkdir(parents=True)
    torch.save(_torch_dict(workloads.adversarial_activations()), d / "adversarial_activations.pt")
    _run("experiments/adversarial_compositional_binding/analyze_adversarial_tda.py", str(tmp_path), monkeypatch)
    out = tmp_path / "tda_adversarial_output"
    with open(out / "summary.json") as f:
        summary = json.load(f)
    assert summary["n_samples_per_condition"] == {"matched": 36, "color_mismatch": 180, "shape_mismatch": 180, "both_mismatch": 324}
    for cond, n in summary["n_samples_per_condition"].items():
        stats = summary["condition_stats"][cond]
        assert len(stats) == 32
        assert list(stats[0]) == ["layer", "n_h1_features", "max_h1_persistence", "max_h0_persistence", "silhouette_img_color",
                                  "silhouette_img_shape", "silhouette_txt_color", "silhouette_txt_shape"]
        cloud = np.load(out / cond / "point_clouds" / "layer_31_cloud.npy")
        assert cloud.shape == (n, 3) and cloud.dtype == np.float32
