import sys, time
sys.path.insert(0, "/root/repo")
import torch
from tda_multimodal_b200 import pipeline, workloads, umap_
from tda_multimodal_b200.pdist import pdist
n = 65536
X = torch.from_numpy(workloads.c5_cloud(n=n, d=4096, seed=5000)).cuda()
def old(X, k=15, row_block=8192):
    parts = []
    for b0 in range(0, n, row_block):
        b1 = min(b0 + row_block, n)
        D = pdist(X[b0:b1][None], X[None], metric="cosine", disconnect=2.0)[0]
        ar = torch.arange(b1 - b0, device=X.device)
        D[ar, b0 + ar] = 0.0
        parts.append(umap_.knn_smooth(D[None], k))
        del D
    return [torch.cat([p[q][0] for p in parts], dim=0) for q in range(4)]
for name, fn in (("python loop", lambda: old(X)), ("tda_knn_fused", lambda: pipeline.knn_row_sharded(X, 15, metric="cosine")), ("python loop", lambda: old(X)), ("tda_knn_fused", lambda: pipeline.knn_row_sharded(X, 15, metric="cosine"))):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    print(f"{name}: {1e3 * (time.perf_counter() - t0):.1f} ms")
a = old(X); b = pipeline.knn_row_sharded(X, 15, metric="cosine")
print("same idx:", torch.equal(a[0], b[0][0]), "same dist:", torch.equal(a[1], b[1][0]))
