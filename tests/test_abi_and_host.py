"""CPU: the C-ABI library loads and exports every symbol include/tda_b200.h declares (no compute calls), host-side
helpers (diagram packing, sharding, stats record), the import shims resolve, and the product fails loudly without a GPU."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "tda_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tda_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from tda_multimodal_b200 import _lib
    syms = _declared_symbols()
    assert len(syms) >= 15
    assert sorted(_lib.PROTOTYPES) == syms, set(syms) ^ set(_lib.PROTOTYPES)
    lib = _lib.lib()
    for s in syms:
        assert hasattr(lib, s), s
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = set(re.findall(r"\bT (tda_[a-z0-9_]+)", out))
    assert set(syms) <= exported
    assert lib.tda_version() >= 1
    # argument validation happens before any CUDA call
    assert lib.tda_rips(None, 0, 0, 1, 0.0, None, None, None, None, 0, None, None, None, 0, 0, None) < 0
    assert b"tda_rips" in lib.tda_last_error()


def test_library_is_sm100a_with_tcgen05_and_tma():
    from tda_multimodal_b200 import _lib
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "UTCHMMA" in sass or "UTCMMA" in sass or re.search(r"UTC\w*MMA", sass)
    assert "UTMALDG" in sass and "LDTM" in sass


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tda_multimodal_b200 import rips, umap_
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rips.ripser(np.zeros((5, 3), np.float32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        umap_.UMAP(n_components=3).fit(np.zeros((5, 3), np.float32))


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "tda_multimodal_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
    for shim in ("umap", "ripser", "persim"):
        src = open(os.path.join(ROOT, "shims", shim, "__init__.py")).read()
        assert "oracle" not in src


def test_shims_resolve():
    code = ("import sys; sys.path.insert(0, %r); import umap, ripser, persim; "
            "print(umap.UMAP.__module__, ripser.ripser.__module__, persim.plot_diagrams.__name__)") % os.path.join(ROOT, "shims")
    out = subprocess.check_output([sys.executable, "-c", code], text=True, cwd="/tmp")
    assert out.split() == ["tda_multimodal_b200.umap_", "tda_multimodal_b200.rips", "plot_diagrams"]


def test_pack_unpack_and_stats_record():
    from tda_multimodal_b200 import pipeline
    rng = np.random.default_rng(0)
    res = []
    for u in range(5):
        h0 = np.c_[np.zeros(4), np.r_[np.sort(rng.uniform(0, 1, 3)), np.inf]]
        h1 = np.sort(rng.uniform(0, 1, (u, 2)), axis=1)
        res.append({"dgms": [h0.astype(np.float32).astype(np.float64), h1.astype(np.float32).astype(np.float64)]})
    counts, payload = pipeline.pack_diagrams(res)
    assert counts.shape == (5, 2) and payload.shape == (counts.sum(), 2)
    back = pipeline.unpack_diagrams(counts, payload)
    for u in range(5):
        assert np.array_equal(back[u][0], res[u]["dgms"][0]) and np.array_equal(back[u][1], res[u]["dgms"][1])
        assert back[u][1].shape == (u, 2) and back[u][1].dtype == np.float64
    rec = pipeline.stats_record(3, res[3]["dgms"])
    assert rec["n_h0_features"] == 1 and rec["n_h1_features"] == 3 and json.dumps(rec)
    assert pipeline.stats_record(0, res[0]["dgms"])["max_h1_persistence"] == 0.0
    assert pipeline.shard_units(10, 1, 4) == [1, 5, 9]


def test_workload_generators_are_seeded():
    from tda_multimodal_b200 import workloads
    a, b = workloads.c3_layer(3, n=50, d=64), workloads.c3_layer(3, n=50, d=64)
    assert a.dtype == np.float32 and a.shape == (50, 64) and np.array_equal(a, b)
    assert not np.array_equal(a, workloads.c3_layer(4, n=50, d=64))
    idx = workloads.c4_resample_indices(0, 200, 4, 100)
    assert idx.shape == (4, 100) and all(len(set(r)) == 100 for r in idx)
    assert (np.diff(idx, axis=1) > 0).all()          # without replacement: ascending sets (the subset front end relies on it)
    assert not (np.diff(workloads.c4_resample_indices(0, 200, 4, 100, replace=True), axis=1) > 0).all()
    c1 = workloads.c1_activations(n_layers=2, d=32)
    assert len(c1) == 48 and sum(v["metadata"]["type"] == "bound" for v in c1.values()) == 36


def test_persim_plot_diagrams_with_stub_matplotlib(monkeypatch):
    """persim.plot_diagrams(dgms, show=False) as the reference calls it (debug_tda_pipeline.py:140).  matplotlib is not in the
    image, so a recording stub stands in for pyplot: what is checked is that the shim accepts ripser-style dgms (float64,
    inf deaths, empty H1) and draws one scatter per dimension plus the diagonal and the infinity line."""
    import types
    calls = {"scatter": [], "plot": [], "legend": 0}

    class Ax:
        def plot(self, *a, **k): calls["plot"].append((a, k))
        def scatter(self, x, y, *a, **k): calls["scatter"].append((np.asarray(x), np.asarray(y), k.get("label")))
        def set_xlabel(self, *_): pass
        def set_ylabel(self, *_): pass
        def set_xlim(self, *_): pass
        def set_ylim(self, *_): pass
        def set_aspect(self, *_): pass
        def set_title(self, *_): pass
        def legend(self, **_): calls["legend"] += 1

    ax = Ax()
    plt = types.ModuleType("matplotlib.pyplot")
    plt.gca = lambda: ax
    plt.show = lambda: None
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = plt
    monkeypatch.setitem(sys.modules, "matplotlib", mpl)
    monkeypatch.setitem(sys.modules, "matplotlib.pyplot", plt)
    monkeypatch.syspath_prepend(os.path.join(ROOT, "shims"))
    sys.modules.pop("persim", None)
    import persim
    h0 = np.array([[0, 0.5], [0, 1.25], [0, np.inf]])
    h1 = np.array([[1.0, 1.5]])
    persim.plot_diagrams([h0, h1], show=False)
    assert len(calls["scatter"]) == 2 and calls["scatter"][0][2] == "$H_{0}$" and calls["legend"] == 1
    assert np.isfinite(calls["scatter"][0][1]).all()          # the infinite bar is drawn on the infinity line
    assert len(calls["plot"]) == 2                            # diagonal + infinity line
    persim.plot_diagrams([h0, np.zeros((0, 2))], show=False)  # empty H1 (happens on the reference's own clouds)
    assert persim.bottleneck(h1, h1) == 0.0
    sys.modules.pop("persim", None)


def test_sweep_grouping_and_cached_index_sets():
    """host logic of the sweep: default group counts (three groups for resident input from 24 layers on, four for host input,
    two from 8 layers on), the option override for the last group, and the C4 index sets (seeded, cached, read-only)."""
    from tda_multimodal_b200 import pipeline, workloads
    assert [pipeline._default_groups(L, False) for L in (1, 7, 8, 23, 24, 32, 64)] == [1, 1, 2, 2, 3, 3, 3]
    assert [pipeline._default_groups(L, True) for L in (4, 8, 24, 32)] == [1, 2, 4, 4]
    assert pipeline.TAIL_OPTIONS == {"rips_cluster": 8}
    a = workloads.c4_resample_indices(3, 2000, 8, 1000)
    b = workloads.c4_resample_indices(3, 2000, 8, 1000)
    assert a is b and not a.flags.writeable and a.shape == (8, 1000)
    assert not np.array_equal(a, workloads.c4_resample_indices(4, 2000, 8, 1000))


def test_bench_algorithmic_work_covers_every_stage():
    """bench.py's roofline table: every stage the library times has an algorithmic-work figure."""
    import importlib.util
    import types
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from tda_multimodal_b200 import _lib
    a = types.SimpleNamespace(points=2000, dim=4096, neighbors=15)
    extra = {"sgd_fired_per_layer": 6.3e6, "reduce_rows_per_layer": 2.2e6, "reduce_heavy_rows_per_layer": 1.6e5}
    for stage in _lib.STAGES:
        kind, work = bench.algorithmic_work(stage, a, 32, extra)
        assert kind in ("hbm", "tensor") and work > 0, stage
    assert bench.union_ms([(0, 2), (1, 3), (5, 6)]) == 4.0
