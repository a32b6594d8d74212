"""CPU, world_size 2, gloo: the N>1 host path -- units sharded u -> rank u mod G, diagrams exchanged with two
all_gathers (counts, padded payload), every rank ends with the full list in unit order."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fake_result(u):
    rng = np.random.default_rng(100 + u)
    h0 = np.c_[np.zeros(6), np.r_[np.sort(rng.uniform(0, 1, 5)), np.inf]].astype(np.float32).astype(np.float64)
    h1 = np.sort(rng.uniform(0, 1, (u % 4, 2)), axis=1).astype(np.float32).astype(np.float64)
    return {"dgms": [h0, h1]}


def _worker(rank, world, port, n_units, ret):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from tda_multimodal_b200 import pipeline
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    units = pipeline.shard_units(n_units, rank, world)
    full = pipeline.gather_diagrams(units, [_fake_result(u) for u in units], n_units)
    ok = all(np.array_equal(full[u][q], _fake_result(u)["dgms"][q]) for u in range(n_units) for q in range(2))
    ret[rank] = bool(ok) and len(full) == n_units
    dist.destroy_process_group()


def _run(n_units):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_units, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret[0] and ret[1]


def test_gather_diagrams_even_and_ragged():
    _run(8)
    _run(5)  # ranks own 3 and 2 units: ragged counts are padded


def test_single_process_passthrough():
    sys.path.insert(0, ROOT)
    from tda_multimodal_b200 import pipeline
    full = pipeline.gather_diagrams([0, 1, 2], [_fake_result(u) for u in range(3)], 3)
    assert all(np.array_equal(full[u][1], _fake_result(u)["dgms"][1]) for u in range(3))


def _bcast_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from tda_multimodal_b200 import pipeline
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)
    raw, emb = torch.randn(37, 16, generator=g), torch.randn(37, 3, generator=g)
    state = pipeline.pack_fitted_state(raw[None], emb[None], 1.577, 0.895, 15) if rank == 0 else None
    got = pipeline.broadcast_fitted_state(state, src=0)
    ok = torch.equal(got["raw_data"], raw) and torch.equal(got["embedding"], emb)
    ok = ok and [float(v) for v in got["scalars"]] == [1.577, 0.895, 15.0, 37.0, 16.0, 3.0]
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_fitted_state_broadcast():
    """analyze_tda_over_layers.py: one fit, many transforms -- the fitted state (training data, embedding, a, b, k) reaches
    every rank unchanged (world size 2, gloo); the transforms themselves are GPU work."""
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_bcast_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret[0] and ret[1]
    sys.path.insert(0, ROOT)
    import torch
    from tda_multimodal_b200 import pipeline
    st = pipeline.pack_fitted_state(torch.zeros(1, 4, 2), torch.zeros(1, 4, 3), 1.0, 1.0, 3)
    assert pipeline.broadcast_fitted_state(st) is st       # single process: passthrough
