"""Import-compatible front door for `from persim import plot_diagrams` (debug_tda_pipeline.py:11,140;
analyze_tda_over_layers.py:6,130; analyze_adversarial_tda.py:13,127).  Plotting is outside the hot path
(SURVEY.md section 8a row a19): this is a thin matplotlib scatter with persim's signature; it accepts the `dgms`
list produced by the ripser shim.  matplotlib is imported lazily (it is not part of the GPU image)."""
import numpy as np

__all__ = ["plot_diagrams", "bottleneck"]


def plot_diagrams(diagrams, plot_only=None, title=None, xy_range=None, labels=None, colormap="default", size=20,
                  ax_color=np.array([0.0, 0.0, 0.0]), diagonal=True, lifetime=False, legend=True, show=False, ax=None):
    import matplotlib.pyplot as plt
    ax = ax or plt.gca()
    if isinstance(diagrams, np.ndarray) and diagrams.ndim == 2:
        diagrams = [diagrams]
    if labels is None:
        labels = ["$H_{{{}}}$".format(i) for i in range(len(diagrams))]
    if plot_only is not None:
        diagrams = [diagrams[i] for i in plot_only]
        labels = [labels[i] for i in plot_only]
    diagrams = [np.asarray(d, dtype=np.float64).reshape(-1, 2) for d in diagrams]
    concat = np.concatenate(diagrams) if diagrams else np.zeros((0, 2))
    finite = concat[np.isfinite(concat).all(axis=1)] if concat.size else concat
    has_inf = bool(concat.size) and bool(np.isinf(concat).any())
    if xy_range is None:
        lo = float(finite.min()) if finite.size else 0.0
        hi = float(finite.max()) if finite.size else 1.0
        buf = (hi - lo) / 5 if hi > lo else 0.2
        x_down, x_up, y_down, y_up = lo - buf / 2, hi + buf, lo - buf / 2, hi + buf
    else:
        x_down, x_up, y_down, y_up = xy_range
    yr = y_up - y_down
    if lifetime:
        y_down, y_up = -yr * 0.05, y_up - y_down
        ax.plot([x_down, x_up], [0, 0], c=ax_color)
    elif diagonal:
        ax.plot([x_down, x_up], [x_down, x_up], "--", c=ax_color)
    b_inf = y_down + yr * 0.95
    if has_inf:
        ax.plot([x_down, x_up], [b_inf, b_inf], "--", c="k", label=r"$\infty$")
    for dgm, label in zip(diagrams, labels):
        d = dgm.copy()
        if lifetime:
            d[:, 1] = d[:, 1] - d[:, 0]
        d[np.isinf(d)] = b_inf
        ax.scatter(d[:, 0], d[:, 1], size, label=label, edgecolor="none")
    ax.set_xlabel("Birth")
    ax.set_ylabel("Lifetime" if lifetime else "Death")
    ax.set_xlim([x_down, x_up])
    ax.set_ylim([y_down, y_up])
    ax.set_aspect("equal", "box")
    if title is not None:
        ax.set_title(title)
    if legend:
        ax.legend(loc="lower right")
    if show:
        plt.show()


def bottleneck(dgm1, dgm2):
    """Bottleneck distance between two diagrams (persim.bottleneck): binary search over candidate costs with a
    bipartite matching feasibility test (Hopcroft-Karp via scipy).  Infinite bars must agree in count and are matched
    by birth order."""
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import maximum_bipartite_matching
    a = np.asarray(dgm1, dtype=np.float64).reshape(-1, 2)
    b = np.asarray(dgm2, dtype=np.float64).reshape(-1, 2)
    ai, bi = np.isinf(a[:, 1]), np.isinf(b[:, 1])
    if ai.sum() != bi.sum():
        return np.inf
    d_inf = float(np.abs(np.sort(a[ai, 0]) - np.sort(b[bi, 0])).max()) if ai.any() else 0.0
    a, b = a[~ai], b[~bi]
    na, nb = len(a), len(b)
    if na + nb == 0:
        return d_inf
    # augmented matching: a_i <-> b_j, a_i <-> diag, diag <-> b_j, diag <-> diag
    da = (a[:, 1] - a[:, 0]) / 2
    db = (b[:, 1] - b[:, 0]) / 2
    C = np.zeros((na + nb, na + nb))
    if na and nb:
        C[:na, :nb] = np.maximum(np.abs(a[:, None, 0] - b[None, :, 0]), np.abs(a[:, None, 1] - b[None, :, 1]))
    C[:na, nb:] = np.inf
    C[na:, :nb] = np.inf
    for i in range(na):
        C[i, nb + i] = da[i]
    for j in range(nb):
        C[na + j, j] = db[j]
    cand = np.unique(C[np.isfinite(C)])
    lo, hi = 0, len(cand) - 1
    while lo < hi:
        mid = (lo + hi) // 2
        m = maximum_bipartite_matching(csr_matrix(C <= cand[mid]), perm_type="column")
        if (m >= 0).all():
            hi = mid
        else:
            lo = mid + 1
    return max(float(cand[lo]), d_inf)
