"""Synthetic inputs of the shapes BASELINE.json names.

The reference's real inputs (all_activations.pt, extract_activations.py:129-132) are git-ignored and need the
Qwen-VL checkpoint, so every measured / tested workload is generated here from seeded numpy generators:
activations emulate LLM statistics as x = s * (Q z + eps) + offset with Q an orthonormal [d, m] frame, z a
low-dimensional latent (torus, clusters, ...) and eps isotropic noise.  Host-side numpy only (no GPU needed).
"""
import numpy as np


def _frame(d, m, rng):
    return np.linalg.qr(rng.standard_normal((d, m)))[0].astype(np.float32)


def _embed(z, d, rng, noise, scale=1.0, offset=0.0):
    Q = _frame(d, z.shape[1], rng)
    X = z.astype(np.float32) @ Q.T
    X += rng.standard_normal(X.shape, dtype=np.float32) * np.float32(noise)
    X *= np.float32(scale)
    if offset:
        X += np.float32(offset)
    return np.ascontiguousarray(X, dtype=np.float32)


def torus_latent(n, rng, jitter=0.02):
    th, ph = rng.uniform(0, 2 * np.pi, (2, n))
    z = np.c_[np.cos(th), np.sin(th), np.cos(ph), np.sin(ph)]
    return z + rng.normal(0, jitter, z.shape)


def c1_activations(seed=1000, n_layers=32, d=4096, metadata=None):
    """C1: the 48-sample 6x6 set of debug_tda_pipeline.py: {id: {'metadata': item, 'activations': {'layer_i': vec}}}
    as numpy vectors (tests wrap them in torch tensors to write all_activations.pt).  Latent = (colour angle,
    shape angle) on S1 x S1; layers differ in how strongly the two factors are mixed / clustered."""
    colors = ["red", "green", "blue", "yellow", "purple", "orange"]
    shapes = ["circle", "square", "triangle", "star", "hexagon", "diamond"]
    if metadata is None:
        metadata = []
        for c in colors:
            for s in shapes:
                metadata.append({"id": f"{c}_{s}", "type": "bound", "color": c, "shape": s, "image_path": f"images/{c}_{s}.png", "prompt": ""})
        for c in colors:
            metadata.append({"id": f"{c}_only", "type": "color_only", "color": c, "shape": None, "image_path": "", "prompt": ""})
        for s in shapes:
            metadata.append({"id": f"{s}_only", "type": "shape_only", "color": None, "shape": s, "image_path": "", "prompt": ""})
    cvals = sorted({m["color"] for m in metadata if m.get("color")})
    svals = sorted({m["shape"] for m in metadata if m.get("shape")})
    out = {m["id"]: {"metadata": m, "activations": {}} for m in metadata}
    for layer in range(n_layers):
        rng = np.random.default_rng(seed + layer)
        mix = layer / max(1, n_layers - 1)
        z = np.zeros((len(metadata), 4))
        for r, m in enumerate(metadata):
            ca = 2 * np.pi * (cvals.index(m["color"]) / len(cvals)) if m.get("color") else rng.uniform(0, 2 * np.pi)
            sa = 2 * np.pi * (svals.index(m["shape"]) / len(svals)) if m.get("shape") else rng.uniform(0, 2 * np.pi)
            z[r] = [np.cos(ca), np.sin(ca), (0.2 + mix) * np.cos(sa), (0.2 + mix) * np.sin(sa)]
        X = _embed(z + rng.normal(0, 0.05, z.shape), d, rng, noise=0.05 / np.sqrt(d) * 4, scale=10.0 * (1 + layer / 8), offset=0.5)
        for r, m in enumerate(metadata):
            out[m["id"]]["activations"][f"layer_{layer}"] = X[r].copy()
    return out


def adversarial_activations(seed=6000, n_layers=32, d=4096):
    """Inputs of experiments/adversarial_compositional_binding/analyze_adversarial_tda.py:34-48: {id: {'metadata': {...},
    'activations': {'layer_i': vec}}} with the four conditions of generate_adversarial_metadata.py:39-108 (36 matched, 180
    colour-mismatch, 180 shape-mismatch, 324 both-mismatch samples).  Latent = image (colour, shape) angles plus a text
    component whose weight grows with the layer."""
    from itertools import product
    colors = ["red", "green", "blue", "yellow", "cyan", "magenta"]
    shapes = ["cube", "sphere", "pyramid", "cone", "torus", "cylinder"]
    meta = []
    for ic, ish in product(colors, shapes):
        base = f"{ic}_{ish}"
        meta.append(dict(id=f"{base}_matched", condition="matched", img_color=ic, img_shape=ish, txt_color=ic, txt_shape=ish))
        meta += [dict(id=f"{base}_color_{tc}", condition="color_mismatch", img_color=ic, img_shape=ish, txt_color=tc, txt_shape=ish)
                 for tc in colors if tc != ic]
        meta += [dict(id=f"{base}_shape_{ts}", condition="shape_mismatch", img_color=ic, img_shape=ish, txt_color=ic, txt_shape=ts)
                 for ts in shapes if ts != ish]
        oc, osh = [c for c in colors if c != ic][:3], [s_ for s_ in shapes if s_ != ish][:3]
        meta += [dict(id=f"{base}_both_{tc}_{ts}", condition="both_mismatch", img_color=ic, img_shape=ish, txt_color=tc, txt_shape=ts)
                 for tc, ts in product(oc, osh)]
    out = {m["id"]: {"metadata": m, "activations": {}} for m in meta}
    ang = lambda vals, v: 2 * np.pi * vals.index(v) / len(vals)
    for layer in range(n_layers):
        rng = np.random.default_rng(seed + layer)
        mix = layer / max(1, n_layers - 1)
        z = np.zeros((len(meta), 8))
        for r, m in enumerate(meta):
            a1, a2 = ang(colors, m["img_color"]), ang(shapes, m["img_shape"])
            b1, b2 = ang(colors, m["txt_color"]), ang(shapes, m["txt_shape"])
            z[r] = [np.cos(a1), np.sin(a1), np.cos(a2), np.sin(a2), mix * np.cos(b1), mix * np.sin(b1), mix * np.cos(b2), mix * np.sin(b2)]
        X = _embed(z + rng.normal(0, 0.05, z.shape), d, rng, noise=0.05 / np.sqrt(d) * 4, scale=8.0 * (1 + layer / 8), offset=0.4)
        for r, m in enumerate(meta):
            out[m["id"]]["activations"][f"layer_{layer}"] = X[r].copy()
    return out


def c2_torus(seed=2000, n=2000, d=4096, sigma=0.02):
    """C2: noisy S1 x S1 torus, n points embedded in d dimensions (ripser on the raw distance matrix)."""
    rng = np.random.default_rng(seed)
    return _embed(torus_latent(n, rng, sigma), d, rng, noise=sigma / np.sqrt(d))


def c3_layer(layer, n=2000, d=4096, seed=3000, n_layers=32, out=None):
    """C3: one layer of the 32 x [2000, 4096] sweep: latent torus blended with 8 Gaussian clusters; the blend,
    scale and common offset vary by layer (early layers toroidal, late layers clustered)."""
    rng = np.random.default_rng(seed + layer)
    frac = layer / max(1, n_layers - 1)
    z_t = torus_latent(n, rng, 0.05)
    centers = rng.normal(0, 1.5, (8, 4))
    lab = rng.integers(0, 8, n)
    z_c = centers[lab] + rng.normal(0, 0.25, (n, 4))
    pick = rng.uniform(size=n) < frac
    z = np.where(pick[:, None], z_c, z_t)
    z = np.c_[z, rng.normal(0, 0.1, (n, 4))]
    X = _embed(z, d, rng, noise=0.02, scale=5.0 * (1.0 + 0.2 * layer), offset=0.3 * (1 + layer % 5))
    if out is not None:
        out[...] = X
        return out
    return X


def c3_layers(n_layers=32, n=2000, d=4096, seed=3000, layers=None, out=None):
    layers = list(range(n_layers)) if layers is None else list(layers)
    if out is None:
        out = np.empty((len(layers), n, d), dtype=np.float32)
    for r, layer in enumerate(layers):
        c3_layer(layer, n, d, seed, n_layers, out=out[r])
    return out


def c4_resample_indices(layer, n_points=2000, n_resamples=256, size=1000, seed=4000, replace=False):
    """C4: bootstrap resamples of one layer's 3-D cloud: [n_resamples, size] int64 indices (without replacement by
    default: duplicate points create zero-length edges; both are supported).  Without replacement every index set is
    ASCENDING: a resample is a set of points, and in the parent's order its edges keep the parent's filtration order, which
    pipeline.bootstrap_rips uses (rips_subsets_launch)."""
    key = (int(layer), int(n_points), int(n_resamples), int(size), int(seed), bool(replace))
    if key not in _C4_INDEX_CACHE:
        out = np.empty((n_resamples, size), dtype=np.int64)
        for r in range(n_resamples):
            rng = np.random.default_rng(seed + 256 * layer + r)
            pick = rng.choice(n_points, size=size, replace=replace)
            out[r] = pick if replace else np.sort(pick)
        out.setflags(write=False)
        if len(_C4_INDEX_CACHE) > 4096:
            _C4_INDEX_CACHE.clear()
        _C4_INDEX_CACHE[key] = out
    return _C4_INDEX_CACHE[key]


_C4_INDEX_CACHE = {}   # the index sets are part of the workload definition (seeded), not of the measured path


def c5_cloud(n=100000, d=4096, seed=5000, latent_dim=20, n_mix=16):
    """C5: one big layer: mixture of n_mix low-dimensional (latent_dim) Gaussian sheets."""
    rng = np.random.default_rng(seed)
    lab = rng.integers(0, n_mix, n)
    centers = rng.normal(0, 2.0, (n_mix, latent_dim))
    z = centers[lab] + rng.normal(0, 0.5, (n, latent_dim))
    Q = _frame(d, latent_dim, rng)
    X = np.empty((n, d), dtype=np.float32)
    step = 8192
    for s in range(0, n, step):
        blk = z[s:s + step].astype(np.float32) @ Q.T
        blk += rng.standard_normal(blk.shape, dtype=np.float32) * np.float32(0.02)
        X[s:s + step] = blk
    return X
