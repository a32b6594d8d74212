import os, sys, time
sys.path.insert(0, "/root/repo")
import torch, numpy as np
from tda_multimodal_b200 import workloads, pipeline, umap_
X = torch.from_numpy(workloads.c3_layers()).cuda()
import functools
orig = umap_.umap_fit_batch
for init in ("spectral", "random"):
    pipeline.umap_fit_batch = functools.partial(orig, init=init)
    for chunks in (1, 2, 4):
        for _ in range(2): pipeline.layer_sweep(X, chunks=chunks)
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(4): pipeline.layer_sweep(X, chunks=chunks)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 4
        print(f"init={init} chunks={chunks}: {dt*1e3:.1f} ms/step", flush=True)
