"""The north-star workload end to end: the 32-layer UMAP + Rips(H0/H1) sweep (config C3) followed by bootstrap resamples of
every layer's 3-D cloud (config C4: R resamples of 1000 points, Rips H0/H1), layers sharded over the ranks, one NCCL gather of
all diagrams at the end.
  python scripts/run_full_target.py [resamples_per_layer=256]
  python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 scripts/run_full_target.py 256"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
from tda_multimodal_b200 import pipeline, workloads

R = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
layers = pipeline.shard_units(32, rank, world)
X = torch.from_numpy(workloads.c3_layers(layers=layers)).cuda()
pipeline.layer_sweep(X[:2])                                   # warm-up (module load, allocator)
if world > 1:
    dist.barrier()
torch.cuda.synchronize(); t0 = time.perf_counter()
out = pipeline.layer_sweep(X)
torch.cuda.synchronize(); t1 = time.perf_counter()
boot = pipeline.bootstrap_rips(out["embedding"], n_resamples=R, size=1000, layer_ids=layers)
torch.cuda.synchronize(); t2 = time.perf_counter()
units = [{"dgms": r["dgms"]} for r in out["results"]] + [{"dgms": r["dgms"]} for per_layer in boot for r in per_layer]
if world > 1:
    counts, payload = pipeline.pack_diagrams(units)
    sizes = torch.tensor([counts.shape[0], payload.shape[0]], device="cuda")
    all_sizes = [torch.empty_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    mc, mp = max(int(s[0]) for s in all_sizes), max(int(s[1]) for s in all_sizes)
    cpad = torch.zeros((mc, 2), dtype=torch.int32, device="cuda"); cpad[:counts.shape[0]] = torch.from_numpy(counts).cuda()
    ppad = torch.zeros((mp, 2), dtype=torch.float32, device="cuda"); ppad[:payload.shape[0]] = torch.from_numpy(payload).cuda()
    gc = [torch.empty_like(cpad) for _ in range(world)]; gp = [torch.empty_like(ppad) for _ in range(world)]
    dist.all_gather(gc, cpad); dist.all_gather(gp, ppad)
    n_diagrams = sum(int(s[0]) for s in all_sizes)
else:
    n_diagrams = len(units)
torch.cuda.synchronize(); t3 = time.perf_counter()
ms = torch.tensor([t1 - t0, t2 - t1, t3 - t2, t3 - t0], device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    a, b, c, tot = ms.tolist()
    print(f"[target] world={world}: 32 layers UMAP+Rips {a*1e3:.0f} ms, 32 x {R} resamples {b*1e3:.0f} ms, gather of {n_diagrams} diagram sets {c*1e3:.0f} ms, "
          f"total {tot:.2f} s = {32 / tot:.1f} layers/s incl. bootstraps ({32 * R / max(b, 1e-9):.0f} resamples/s)", flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
