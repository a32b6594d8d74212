"""CPU: pins the Rips oracle on the reference's own shipped artefacts:
tda-output/point_clouds_3d/layer_i_cloud.npy -> tda-output/summary_stats.json, 32 layers."""
import numpy as np
import pytest

from oracle import rips as orips
from tests.helpers import load_ref_rips_golden, reference_stats, torus3d, circle2d, same_diagram


@pytest.mark.parametrize("sklearn_dm", [False, True])
def test_oracle_reproduces_reference_summary_stats(sklearn_dm):
    clouds, stats = load_ref_rips_golden()
    assert clouds.shape == (32, 36, 3) and len(stats) == 32
    for i in range(32):
        res = orips.ripser(clouds[i], maxdim=1, _sklearn_dm=sklearn_dm)
        got = reference_stats(res["dgms"])
        for key in ("n_h1_features", "max_h1_persistence", "all_h1_persistence_values", "n_h0_features", "max_h0_persistence"):
            assert got[key] == stats[i][key], (i, key)  # bit-exact, order-sensitive for the H1 list
        assert res["dgms"][0].dtype == np.float64 and res["dgms"][0].shape == (36, 2)


def test_oracle_known_answers():
    # 4 corners of the unit square: one H1 bar [1, sqrt(2))
    sq = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], dtype=np.float32)
    d = orips.ripser(sq, maxdim=1)["dgms"]
    assert d[1].shape == (1, 2) and d[1][0, 0] == 1.0 and d[1][0, 1] == np.float32(np.sqrt(np.float32(2.0)))
    assert np.isinf(d[0][-1, 1]) and (d[0][:-1, 1] == 1.0).all()
    # points on a line: no H1
    line = np.c_[np.arange(10, dtype=np.float32) ** 1.5, np.zeros(10, dtype=np.float32)]
    assert orips.ripser(line, maxdim=1)["dgms"][1].shape == (0, 2)
    # a noisy circle: exactly one dominant H1 bar
    rng = np.random.default_rng(3)
    d1 = orips.ripser(circle2d(200, rng), maxdim=1)["dgms"][1]
    pers = np.sort(d1[:, 1] - d1[:, 0])[::-1]
    assert pers[0] > 1.0 and pers[1] < 0.3
    # duplicate points: zero-length edges give no H0 row
    dup = np.array([[0, 0], [0, 0], [1, 0]], dtype=np.float32)
    d0 = orips.ripser(dup, maxdim=0)["dgms"][0]
    assert d0.shape == (2, 2) and d0[0, 1] == 1.0


def test_oracle_invariants():
    rng = np.random.default_rng(11)
    X = torus3d(150, rng)
    res = orips.ripser(X, maxdim=1)
    d0, d1 = res["dgms"]
    assert d0.shape[0] == 150 and np.isinf(d0[-1, 1]) and np.all(np.diff(d0[:-1, 1]) >= 0)
    assert np.all(np.diff(d1[:, 0]) <= 0)  # H1 rows in birth-descending order
    perm = rng.permutation(150)
    res2 = orips.ripser(X[perm], maxdim=1)
    assert same_diagram(res["dgms"][1], res2["dgms"][1]) and same_diagram(d0, res2["dgms"][0])
    # H0 deaths are the MST weights
    from scipy.sparse.csgraph import minimum_spanning_tree
    dm = orips.euclidean_dm_f32(X).astype(np.float64)
    mst = np.sort(minimum_spanning_tree(dm).data)
    assert np.allclose(mst, d0[:-1, 1])
    # torus (R=3, r=1): the big circle is the dominant H1 class
    pers = np.sort(d1[:, 1] - d1[:, 0])[::-1]
    assert pers[0] > 2 * pers[1]


def test_oracle_maxdim2_sphere():
    rng = np.random.default_rng(5)
    v = rng.normal(size=(60, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    d = orips.ripser(v.astype(np.float32), maxdim=2)["dgms"]
    pers2 = np.sort(d[2][:, 1] - d[2][:, 0])[::-1]
    assert len(pers2) >= 1 and pers2[0] > 0.3 and (len(pers2) == 1 or pers2[0] > 2 * pers2[1])


def test_c2_full_size_golden_agrees_with_the_logged_gpu_run():
    """Config C2 at its full size (2000-point torus in 4096-d, maxdim=2): the oracle's diagrams (tests/golden/c2_torus_n2000_dgms.npz,
    650 s in the oracle's lean mode) against the GPU run recorded in profiles/r02_c2_n2000.log earlier in the round -- same number
    of edges, H1 and H2 rows, same top persistences to the three decimals the log holds.  (The GPU suite compares the diagrams
    themselves by bottleneck distance: tests/test_rips_h2_gpu.py.)"""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gold = np.load(os.path.join(root, "tests", "golden", "c2_torus_n2000_dgms.npz"))
    line = [l for l in open(os.path.join(root, "profiles", "r02_c2_n2000.log")) if l.startswith("C2 n=2000")][0]
    m = re.search(r"num_edges (\d+); H1 rows (\d+) top \[([^\]]*)\]; H2 rows (\d+) top \[([^\]]*)\]", line)
    assert m, line
    assert int(m.group(1)) == int(gold["num_edges"]) and int(m.group(2)) == len(gold["h1"]) and int(m.group(4)) == len(gold["h2"])
    for name, txt in (("h1", m.group(3)), ("h2", m.group(5))):
        logged = np.array([float(t) for t in txt.split()])
        pers = np.sort(gold[name][:, 1] - gold[name][:, 0])[::-1][:len(logged)]
        assert np.allclose(pers, logged, atol=6e-4), (name, pers, logged)
    assert gold["h0"].shape == (2000, 2) and np.isinf(gold["h0"][-1, 1])
