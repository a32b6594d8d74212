"""GPU: the substitute-then-verify variant of the sweep reducer (TDA_RIPS_REDUCER=verify; DESIGN.md section 6, CPU model in
oracle/rips_propagate_model.cpp) must give the oracle's diagrams and simplex pairs bit for bit, like the default reducer.
Written at the end of round 1 with no GPU minutes left: compiled, never run -- skipped until TDA_TEST_UNVALIDATED=1."""
import os

import numpy as np
import pytest

from tests.helpers import blobs3d, circle2d, load_ref_rips_golden, torus3d

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("TDA_TEST_UNVALIDATED") != "1", reason="not yet run on a GPU (set TDA_TEST_UNVALIDATED=1)")]


@pytest.mark.parametrize("gen,n,seed", [(torus3d, 60, 0), (circle2d, 150, 1), (blobs3d, 400, 2), (torus3d, 1000, 3), (blobs3d, 2000, 4)])
def test_verify_reducer_matches_oracle(monkeypatch, gen, n, seed):
    import torch
    from oracle import rips as orips
    from tda_multimodal_b200 import rips
    monkeypatch.setenv("TDA_RIPS_REDUCER", "verify")
    X = gen(n, np.random.default_rng(seed))
    want = orips.ripser(X, maxdim=1)["dgms"]
    got = rips.ripser(X, maxdim=1)["dgms"]
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    torch.cuda.synchronize()


def test_verify_reducer_on_reference_clouds_batched(monkeypatch):
    import torch
    from oracle import rips as orips
    from tda_multimodal_b200 import rips
    monkeypatch.setenv("TDA_RIPS_REDUCER", "verify")
    clouds, _ = load_ref_rips_golden()
    dm = rips.pdist_lowdim(torch.from_numpy(clouds.astype(np.float32)).cuda())
    res = rips.rips_batch(dm, maxdim=1)
    for i in range(len(clouds)):
        want = orips.ripser(clouds[i], maxdim=1)["dgms"]
        assert np.array_equal(res[i]["dgms"][0], want[0]) and np.array_equal(res[i]["dgms"][1], want[1])
