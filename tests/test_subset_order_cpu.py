"""CPU models (numpy, test infrastructure) of the two claims behind the bootstrap front end and the one-launch H0
(DESIGN.md sections 2.6 and 2.9; the kernels are subset_count/scatter_kernel and boruvka_chunked_kernel in csrc/rips.cu):

1. a subset of a cloud given by ASCENDING parent indices keeps the parent's filtration order of its edges, ties included, so the
   filtration rank of a subset edge is the number of subset edges before it in the parent's sorted list;
2. Boruvka rounds restricted to successive chunks of the sorted edge list (lowest rank of the chunk leaving each component,
   repeated until the chunk holds no joining edge) pick exactly the edges Kruskal / ripser's union-find sweep picks.
"""
import numpy as np
import pytest


def _sorted_edges(pts):
    """ripser's edge order: length ascending (float32, as tda_pdist_lowdim rounds), edge index C(i,2)+j descending."""
    n = len(pts)
    i, j = np.tril_indices(n, -1)                     # i > j
    d = np.sqrt(((pts[i].astype(np.float64) - pts[j].astype(np.float64)) ** 2).sum(1).astype(np.float32)).astype(np.float32)
    index = i * (i - 1) // 2 + j
    order = np.lexsort((-index, d))
    return i[order], j[order], d[order]


def _clouds(rng):
    lattice = np.stack(np.meshgrid(np.arange(5), np.arange(5), np.arange(4), indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    lattice = np.concatenate([lattice, lattice[:15]])[rng.permutation(115)]          # ties and duplicate points
    blobs = np.concatenate([rng.normal(size=(40, 3)) * 0.05 + c for c in ((0, 0, 0), (6, 0, 0), (0, 5, 3))]).astype(np.float32)
    return [rng.normal(size=(120, 3)).astype(np.float32), lattice, blobs]


@pytest.mark.parametrize("which", [0, 1, 2])
def test_subset_ranks_are_a_prefix_count_over_the_parents_order(which):
    rng = np.random.default_rng(900 + which)
    pts = _clouds(rng)[which]
    n = len(pts)
    pi, pj, pd = _sorted_edges(pts)
    for m in (n, n // 2, 7):
        idx = np.sort(rng.choice(n, size=m, replace=False))
        mask = np.full(n, -1)
        mask[idx] = np.arange(m)
        flag = (mask[pi] >= 0) & (mask[pj] >= 0)
        rank_sub = np.cumsum(flag) - 1                 # exclusive prefix count at the flagged positions
        got_i, got_j, got_d = mask[pi[flag]], mask[pj[flag]], pd[flag]
        assert np.array_equal(rank_sub[flag], np.arange(m * (m - 1) // 2))
        want_i, want_j, want_d = _sorted_edges(pts[idx])
        assert np.array_equal(got_i, want_i) and np.array_equal(got_j, want_j) and np.array_equal(got_d, want_d)
    # ... and NOT for an unordered index set on a cloud with ties: the tie-break follows the relabelling
    if which == 1:
        idx = rng.permutation(n)[: n // 2]
        mask = np.full(n, -1)
        mask[idx] = np.arange(len(idx))
        flag = (mask[pi] >= 0) & (mask[pj] >= 0)
        a, b = mask[pi[flag]], mask[pj[flag]]
        got = np.stack([np.maximum(a, b), np.minimum(a, b)], 1)
        wi, wj, _ = _sorted_edges(pts[idx])
        assert not np.array_equal(got, np.stack([wi, wj], 1))


def _kruskal(n, ei, ej, T):
    parent = list(range(n))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x
    picked = []
    for r in range(T):
        a, b = find(int(ei[r])), find(int(ej[r]))
        if a != b:
            parent[a] = b
            picked.append(r)
    return picked


def _boruvka_by_chunks(n, ei, ej, T, chunk):
    """boruvka_chunked_kernel, sequentially: same candidate rule, same hooking (mutual picks: the smaller label is the root)."""
    comp = np.arange(n)
    picked = set()
    c0, length = 0, chunk
    while c0 < T and len(picked) < n - 1:
        c1 = min(T, c0 + length)
        best = {}
        for r in range(c0, c1):
            cu, cv = comp[ei[r]], comp[ej[r]]
            if cu != cv:
                for c in (cu, cv):
                    if r < best.get(c, 1 << 60):
                        best[c] = r
        if not best:
            c0, length = c1, length * 2
            continue
        parent = np.arange(n)
        for c, r in best.items():
            cu, cv = comp[ei[r]], comp[ej[r]]
            parent[c] = cv if cu == c else cu
            picked.add(r)
        for c in best:
            q = parent[c]
            if q != c and parent[q] == c and c < q:
                parent[c] = c
        while True:
            nxt = parent[parent]
            if np.array_equal(nxt, parent):
                break
            parent = nxt
        comp = parent[comp]
    return sorted(picked)


@pytest.mark.parametrize("which", [0, 1, 2])
@pytest.mark.parametrize("chunk", [16, 257, 1 << 20])
def test_boruvka_by_chunks_is_kruskal(which, chunk):
    rng = np.random.default_rng(950 + which)
    pts = _clouds(rng)[which]
    n = len(pts)
    ei, ej, d = _sorted_edges(pts)
    for T in (len(d), int(np.searchsorted(d, 1.0))):       # whole filtration / a threshold that leaves several components
        assert _boruvka_by_chunks(n, ei, ej, T, chunk) == _kruskal(n, ei, ej, T)
