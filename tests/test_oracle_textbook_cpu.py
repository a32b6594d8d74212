"""CPU: the Rips oracle against an INDEPENDENT algorithm on small clouds.

oracle/rips_oracle.cpp restates Ripser (implicit coboundary reduction in cohomology, emergent pairs, clearing).  Here the same
diagrams are computed the textbook way -- every simplex of the filtration up to dimension 3 written out, the boundary matrix over
Z/2 reduced column by column in filtration order (Edelsbrunner / Letscher / Zomorodian) -- which shares nothing with the oracle but
the float32 distance matrix.  The diagrams (multisets of (birth, death) with birth < death, essential classes as (birth, inf)) must
be equal bit for bit in H0, H1 and H2, with the enclosing-radius threshold, with a finite threshold (essential H1 / H2 classes),
and on clouds with tied distances and duplicate points.  This pins the H2 leg of the oracle (config C2), for which the reference
ships no golden vectors, beyond Betti numbers."""
import itertools

import numpy as np
import pytest

from oracle import rips as orips


def textbook_diagrams(dm, maxdim, thresh):
    """Standard persistence algorithm on the explicit Rips filtration of the float32 matrix `dm` up to `thresh` (inclusive)."""
    n = dm.shape[0]
    simplices = []   # (value, dim, vertices)
    for q in range(maxdim + 2):
        for vs in itertools.combinations(range(n), q + 1):
            val = np.float32(0.0)
            for a, b in itertools.combinations(vs, 2):
                val = max(val, dm[a, b])
            if val <= thresh:
                simplices.append((float(val), q, vs))
    simplices.sort(key=lambda s: (s[0], s[1], s[2]))   # faces come before cofaces: a coface has value >= and dimension >
    index = {s[2]: i for i, s in enumerate(simplices)}
    low_owner = {}                # pivot row -> reduced column (as a Python int bit set)
    paired_birth = set()
    positive = []
    pairs = [[] for _ in range(maxdim + 1)]
    for j, (val, q, vs) in enumerate(simplices):
        col = 0
        if q > 0:
            for k in range(q + 1):
                col ^= 1 << index[vs[:k] + vs[k + 1:]]
        while col:
            low = col.bit_length() - 1
            other = low_owner.get(low)
            if other is None:
                break
            col ^= other
        if col:
            low = col.bit_length() - 1
            low_owner[low] = col
            paired_birth.add(low)
            b, bq = simplices[low][0], simplices[low][1]
            if bq <= maxdim and b < val:
                pairs[bq].append((b, val))
        else:
            positive.append(j)
    for j in positive:
        if j not in paired_birth and simplices[j][1] <= maxdim:
            pairs[simplices[j][1]].append((simplices[j][0], np.inf))
    return [np.array(sorted(p), dtype=np.float64).reshape(-1, 2) for p in pairs]


def enclosing_radius(dm):
    return np.float32(dm.max(axis=1).min())


def sorted_rows(d):
    d = np.asarray(d, dtype=np.float64).reshape(-1, 2)
    return d[np.lexsort((d[:, 1], d[:, 0]))]


def _clouds():
    rng = np.random.default_rng(77)
    out = {"gauss3d": rng.normal(size=(22, 3)).astype(np.float32)}
    th = rng.uniform(0, 2 * np.pi, 20)
    out["circle"] = (np.c_[np.cos(th), np.sin(th)] + rng.normal(0, 0.05, (20, 2))).astype(np.float32)
    v = rng.normal(size=(24, 3))
    out["sphere"] = (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)
    out["grid_ties"] = np.array([[i, j] for i in range(5) for j in range(4)], dtype=np.float32)          # many equal distances
    out["cube_ties"] = np.array(list(itertools.product((0, 1), repeat=3)) + [[0.5, 0.5, 0.5]], dtype=np.float32)
    dup = rng.normal(size=(14, 3)).astype(np.float32)
    out["duplicates"] = np.concatenate([dup, dup[:4]])                                                     # zero-length edges
    out["octahedron"] = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], dtype=np.float32)
    return out


CLOUDS = _clouds()


@pytest.mark.parametrize("apparent", [False, True])
@pytest.mark.parametrize("name", sorted(CLOUDS))
def test_oracle_equals_textbook_reduction_enclosing_radius(name, apparent):
    X = CLOUDS[name]
    dm = orips.euclidean_dm_f32(X)
    got = orips.rips_dm(dm, maxdim=2, apparent=apparent)
    want = textbook_diagrams(dm, 2, enclosing_radius(dm))
    assert got["thresh"] == float(enclosing_radius(dm))
    for q in range(3):
        g = sorted_rows(got["dgms"][q])
        if q == 0:   # one essential component: ripser reports it as (0, inf); the textbook complex holds it as the only unpaired vertex
            assert np.isinf(g[-1, 1])
        assert np.array_equal(g, want[q]), (name, q, g.shape, want[q].shape)


@pytest.mark.parametrize("name,frac", [("gauss3d", 0.6), ("circle", 0.5), ("sphere", 0.7), ("grid_ties", 0.5), ("octahedron", 0.8)])
def test_oracle_equals_textbook_reduction_finite_threshold(name, frac):
    """A threshold below the enclosing radius leaves essential H1 / H2 classes: rows (birth, inf)."""
    X = CLOUDS[name]
    dm = orips.euclidean_dm_f32(X)
    thresh = np.float32(frac * enclosing_radius(dm))
    got = orips.rips_dm(dm, maxdim=2, thresh=float(thresh))
    want = textbook_diagrams(dm, 2, thresh)
    for q in range(3):
        assert np.array_equal(sorted_rows(got["dgms"][q]), want[q]), (name, q)
    if name == "octahedron":   # the hollow octahedron: one essential 2-class once the 12 edges are in and the 3 diagonals are not
        assert want[2].shape == (1, 2) and np.isinf(want[2][0, 1])


@pytest.mark.parametrize("seed", range(12))
def test_oracle_equals_textbook_reduction_random_sweep(seed):
    """Random clouds, half of them on a small integer lattice (heavy ties, duplicates), random dimension and threshold."""
    rng = np.random.default_rng(1000 + seed)
    n, dim = int(rng.integers(8, 27)), int(rng.integers(2, 5))
    X = (rng.integers(0, 4, (n, dim)) if seed % 2 else rng.normal(size=(n, dim))).astype(np.float32)
    dm = orips.euclidean_dm_f32(X)
    enc = enclosing_radius(dm)
    thresh = enc if seed % 3 == 0 else np.float32(rng.uniform(0.4, 1.0) * enc)
    want = textbook_diagrams(dm, 2, thresh)
    for apparent in (False, True):
        got = orips.rips_dm(dm, maxdim=2, thresh=float(thresh), apparent=apparent)
        for q in range(3):
            assert np.array_equal(sorted_rows(got["dgms"][q]), want[q]), (seed, n, dim, q, apparent)


def test_apparent_pair_shortcut_changes_nothing():
    """The oracle's optional apparent-pair shortcut (Ripser 1.2's; what makes config C2 fit in memory at n = 2000) gives the same rows
    in the same order with the same birth / death simplices as the plain reduction: on the reference's 32 shipped clouds, on lattice
    clouds with ties and duplicates under a finite threshold, and on the committed C2 golden at n = 600 (made without it)."""
    import os
    from tests.helpers import load_ref_rips_golden
    from tda_multimodal_b200 import workloads
    clouds, _ = load_ref_rips_golden()
    cases = [(orips.euclidean_dm_f32(c), np.inf) for c in clouds]
    rng = np.random.default_rng(9)
    for t in range(12):
        n = int(rng.integers(10, 90))
        X = (rng.integers(0, 5, (n, 3)) if t % 2 else rng.normal(size=(n, 3))).astype(np.float32)
        dm = orips.euclidean_dm_f32(X)
        cases.append((dm, np.inf if t % 3 else float(0.7 * enclosing_radius(dm))))
    for dm, thresh in cases:
        a = orips.rips_dm(dm, maxdim=2, thresh=thresh, with_simplices=True)
        b = orips.rips_dm(dm, maxdim=2, thresh=thresh, with_simplices=True, apparent=True)
        for q in range(3):
            assert np.array_equal(a["dgms"][q], b["dgms"][q]) and np.array_equal(a["simplices"][q], b["simplices"][q]), q
    X = workloads.c2_torus(n=600).astype(np.float64)
    sq = (X * X).sum(1)
    D = np.sqrt(np.maximum(sq[:, None] + sq[None] - 2.0 * X @ X.T, 0.0))
    np.fill_diagonal(D, 0.0)
    r = orips.rips_dm(D.astype(np.float32), maxdim=2, apparent=True)
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c2_torus_n600_dgms.npz"))
    for q, name in enumerate(("h0", "h1", "h2")):
        assert np.array_equal(r["dgms"][q], gold[name]), name


def test_working_column_compaction_changes_nothing():
    """The oracle removes cancelling pairs from its working columns once they hold 2^25 entries (config C2 at n = 2000 would need tens
    of GB otherwise), and in the lean mode keeps only the cofacets inside a diameter window in the heap, re-enumerating when the window
    moves.  With the bound lowered to 16 entries and the window to one edge rank both mechanisms run all the time: same rows, order
    and simplices."""
    from tests.helpers import torus3d
    rng = np.random.default_rng(21)
    cases = [orips.euclidean_dm_f32(torus3d(120, rng)), orips.euclidean_dm_f32(rng.integers(0, 5, (70, 3)).astype(np.float32)),
             orips.euclidean_dm_f32(rng.normal(size=(90, 3)).astype(np.float32))]
    for dm in cases:
        for apparent in (False, True):
            a = orips.rips_dm(dm, maxdim=2, with_simplices=True, with_stats=True, apparent=apparent)
            try:
                orips.set_compact(16)
                orips.set_window(1, 1 << 40)      # lean mode: the diameter window of the working coboundary starts at ONE edge rank
                b = orips.rips_dm(dm, maxdim=2, with_simplices=True, with_stats=True, apparent=apparent)
            finally:
                orips.set_compact()
                orips.set_window()
            assert a["stats"][1]["additions"] > 0
            for q in range(3):
                assert np.array_equal(a["dgms"][q], b["dgms"][q]) and np.array_equal(a["simplices"][q], b["simplices"][q]), q


def test_lean_mode_reproduces_the_c2_golden_at_n1000():
    """tests/golden/c2_torus_n1000_dgms.npz was made by the plain reduction (5 min, 19 GB); the lean mode (apparent pairs + windowed
    working column, 20 s, 0.5 GB) must give the same three arrays bit for bit -- the n = 2000 golden exists only in that mode."""
    import os
    from tda_multimodal_b200 import workloads
    X = workloads.c2_torus(n=1000).astype(np.float64)
    sq = (X * X).sum(1)
    D = np.sqrt(np.maximum(sq[:, None] + sq[None] - 2.0 * X @ X.T, 0.0))
    np.fill_diagonal(D, 0.0)
    r = orips.rips_dm(D.astype(np.float32), maxdim=2, apparent=True)
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c2_torus_n1000_dgms.npz"))
    for q, name in enumerate(("h0", "h1", "h2")):
        assert np.array_equal(r["dgms"][q], gold[name]), name
