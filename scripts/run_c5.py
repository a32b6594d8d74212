"""Config C5 of BASELINE.json at a chosen size: one large cloud, row-sharded exact kNN over the ranks (NCCL all_gather of the
[n/G, k] blocks), then (rank 0) UMAP 3-D from that kNN and Rips H1 on a landmark subsample.
  python scripts/run_c5.py [n] [landmarks]                      (1 GPU)
  python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 scripts/run_c5.py [n] [landmarks]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
from tda_multimodal_b200 import pipeline, umap_, rips, workloads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
n_land = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
X = torch.from_numpy(workloads.c5_cloud(n=n, d=4096, seed=5000)).cuda()      # replicated (same seed on every rank)
for rep in range(2):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    knn = pipeline.knn_row_sharded(X, 15, metric="cosine")
    torch.cuda.synchronize(); t_knn = time.perf_counter() - t0
if rank == 0:
    flops = 2.0 * n * n * 4096
    print(f"[C5] n={n} world={world}: row-sharded pdist+kNN {t_knn*1e3:.1f} ms  ({flops / t_knn / 1e12:.1f} useful TFLOP/s aggregate, "
          f"{4.0 * n * n / t_knn / 1e9:.0f} GB/s of distance matrix consumed without being stored)", flush=True)
    t0 = time.perf_counter()
    Y = umap_.umap_fit_batch(X[None], n_neighbors=15, n_components=3, metric="cosine", random_state=42, knn=knn)[0]
    torch.cuda.synchronize(); t_umap = time.perf_counter() - t0
    t0 = time.perf_counter()
    r = rips.ripser(Y, maxdim=1, n_perm=n_land)
    t_rips = time.perf_counter() - t0
    d1 = r["dgms"][1]
    print(f"[C5] UMAP (graph, spectral init, SGD) {t_umap*1e3:.0f} ms; landmark Rips ({n_land} of {n}) {t_rips*1e3:.0f} ms; "
          f"H1 rows {len(d1)}, r_cover {r['r_cover']:.3f}", flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
