"""CPU model of the selection inside knn_smooth_block_kernel (csrc/umap.cu): chunks of 128 x loads columns, the k-th smallest
of the 32 lane minima as the candidate bound, candidates pushed as 64-bit (ordered distance bits, column) keys, warp bitonic
sort + bitonic merges into the best-k list.  Checked against numpy's stable argsort (umap-learn's fast_knn_indices) on random
rows with ties, many chunks and small chunks -- the GPU tests (tests/test_umap_gpu.py) check the kernel itself."""
import numpy as np
MAXK = np.uint64(0xFFFFFFFFFFFFFFFF)
def key(d, j):
    d = np.float32(d) + np.float32(0.0)
    b = np.frombuffer(np.float32(d).tobytes(), dtype=np.uint32)[0]
    o = (~b & np.uint32(0xFFFFFFFF)) if (b & np.uint32(0x80000000)) else (b | np.uint32(0x80000000))
    return (np.uint64(o) << np.uint64(32)) | np.uint64(j)
def key_dist(k):
    o = np.uint32(k >> np.uint64(32))
    b = (o ^ np.uint32(0x80000000)) if (o & np.uint32(0x80000000)) else (~o & np.uint32(0xFFFFFFFF))
    return np.frombuffer(np.uint32(b).tobytes(), dtype=np.float32)[0]
def warp_sort_asc(v):
    v = v.copy(); lanes = np.arange(32)
    k2 = 2
    while k2 <= 32:
        j = k2 >> 1
        while j > 0:
            o = v[lanes ^ j]
            keep_min = ((lanes & j) == 0) == ((lanes & k2) == 0)
            mn = np.minimum(v, o); mx = np.maximum(v, o)
            v = np.where(keep_min, mn, mx)
            j >>= 1
        k2 <<= 1
    return v
def warp_merge_bitonic(v):
    v = v.copy(); lanes = np.arange(32)
    j = 16
    while j > 0:
        o = v[lanes ^ j]
        v = np.where((lanes & j) == 0, np.minimum(v, o), np.maximum(v, o))
        j >>= 1
    return v
def select_row(row, k, loads, cand_cap=64):
    m = len(row); lanes = np.arange(32)
    chunk = 128 * loads
    best = np.full(32, MAXK, dtype=np.uint64); tau = np.float32(np.inf)
    buf = []
    n_merge_nonempty = 0
    for c0 in range(0, m, chunk):
        # lane l, load u, comp c -> e = c0 + (u*32+l)*4 + c
        lmin = np.full(32, np.inf, dtype=np.float32)
        elems = []
        for u in range(loads):
            for l in range(32):
                for c in range(4):
                    e = c0 + (u * 32 + l) * 4 + c
                    if e < m:
                        lmin[l] = min(lmin[l], row[e]); elems.append((u, l, c, e))
        sm = warp_sort_asc(lmin)
        tau = min(tau, sm[k - 1])
        for (u, l, c, e) in elems:
            if row[e] <= tau: buf.append(key(row[e], e))
        cnt = len(buf)
        if cnt > cand_cap: return None, 'overflow'
        last = c0 + chunk >= m
        if cnt >= 32 or last:
            a = np.full(32, MAXK, dtype=np.uint64); a[:min(cnt, 32)] = buf[:32]
            a = warp_sort_asc(a)
            if cnt > 32:
                b = np.full(32, MAXK, dtype=np.uint64); b[:cnt - 32] = buf[32:]
                b = warp_sort_asc(b)
                br = b[31 - lanes]
                a = warp_merge_bitonic(np.minimum(a, br))
            if (best != MAXK).any(): n_merge_nonempty += 1
            rev = best[31 - lanes]
            w = warp_merge_bitonic(np.where(lanes < 16, a, rev))
            best = np.where(lanes < k, w, MAXK)
            kth = best[k - 1]
            if kth != MAXK: tau = min(tau, key_dist(kth))
            buf = []
    return best[:k], n_merge_nonempty
def test_selection_model_matches_stable_argsort():
  rng = np.random.default_rng(0)
  tot = 0; ovf = 0; nonempty = 0
  for trial in range(120):
      m = int(rng.choice([36, 130, 500, 1000, 1028, 2000, 3001, 5000, 9000]))
      k = int(rng.choice([2, 6, 10, 15, 16]))
      loads = int(rng.choice([8, 16])) if m < 4000 else 8
      loads = 2 if trial % 5 == 0 else loads          # small chunks: many merges into a non-empty list
      row = rng.random(m, dtype=np.float32)
      if trial % 7 == 0: row[rng.integers(0, m, m // 10)] = row[0]      # ties
      if trial % 11 == 0: row = np.round(row * 8) / 8                    # massive ties -> overflow path
      got, info = select_row(row, k, loads)
      if got is None: ovf += 1; continue
      nonempty += info
      order = np.argsort(row, kind='stable')[:k]
      want = np.array([key(row[j], j) for j in order], dtype=np.uint64)
      assert np.array_equal(got, want), (trial, m, k, loads)
      tot += 1
  assert tot > 80 and nonempty > 20 and ovf > 0, (tot, ovf, nonempty)

