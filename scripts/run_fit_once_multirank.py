"""analyze_tda_over_layers.py's flow (fit UMAP on ONE layer, transform every layer, Rips on each embedding; :56-77) over the ranks of
a node: pipeline.fit_once_transform_many under NCCL -- rank 0 fits, the fitted state is broadcast (2 broadcasts), transforms and Rips
shard by layer, the diagrams are gathered (2 all_gathers).
  python scripts/run_fit_once_multirank.py out.npz                                                     (1 GPU: writes the diagrams)
  python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 scripts/run_fit_once_multirank.py out.npz
      (G ranks: every layer's gathered diagrams must equal the single-rank file bit for bit -- the kernels are deterministic)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
from tda_multimodal_b200 import pipeline, workloads

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/fit_once_w1.npz"
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
# the reference's shape (36 bound samples x 4096-d per layer, k = N // 2 = 18) and a larger one (32 layers x 400 points, k = 15)
cases = []
acts = workloads.c1_activations()
ids = [i for i, v in acts.items() if v["metadata"]["type"] == "bound"]
X1 = np.stack([np.stack([acts[i]["activations"][f"layer_{l}"] for i in ids]) for l in range(32)]).astype(np.float32)
cases.append(("c1_36x4096_k18", X1, 18))
cases.append(("c3_400x256_k15", workloads.c3_layers(n_layers=32, n=400, d=256), 15))
out = {}
for name, Xh, k in cases:
    X = torch.from_numpy(Xh).cuda()
    for rep in range(2):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        emb, dgms = pipeline.fit_once_transform_many(X, n_neighbors=k)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    assert all(d is not None for d in dgms) and len(dgms) == 32
    assert sorted(emb) == pipeline.shard_units(32, rank, world)
    for l, d in enumerate(dgms):
        out[f"{name}_L{l}_h0"] = d[0]; out[f"{name}_L{l}_h1"] = d[1]
    if rank == 0:
        print(f"[fit-once] {name}: world={world}, 32 layers (fit on the last, 31 transforms + 32 Rips), {dt*1e3:.1f} ms, "
              f"H1 rows per layer {[len(d[1]) for d in dgms][:8]}...", flush=True)
if rank == 0:
    if world == 1:
        np.savez(path, **out)
        print(f"[fit-once] wrote {len(out)} diagrams to {path}")
    else:
        ref = np.load(path)
        bad = [k for k in out if not np.array_equal(out[k], ref[k])]
        print(f"[fit-once] world={world}: {len(out) - len(bad)} of {len(out)} gathered diagrams equal the single-rank run bit for bit" + (f"; MISMATCH {bad[:5]}" if bad else ""))
        assert not bad
if world > 1:
    dist.barrier(); dist.destroy_process_group()
