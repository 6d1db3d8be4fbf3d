"""Two torch.distributed ranks (gloo rendezvous, both on cuda:0) each Fit their share of the chains: the R-hat tables
are computed from all-gathered split-half moments and must equal a single-process Fit over all chains; the chains
themselves are identical because the RNG is keyed by the global chain id."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _data():
    rng = np.random.default_rng(3)
    V, n = 7, 50
    q = V * (V + 1) // 2
    X = rng.normal(size=(n, q))
    y = 2.0 + X[:, :4].sum(axis=1) + rng.normal(size=n)
    return X, y


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from conftest import load_package
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bnr = load_package()
    X, y = _data()
    res = bnr.Fit(X, y, 3, nburn=60, nsamples=40, num_chains=3, seed=17, x_transform=False, filename=None,
                  psrf_cutoff=1e9, device=0, return_state="gamma_xi")
    q.put((rank, np.asarray(res.rhatγ.γ), np.asarray(res.rhatξ.ξ), res.state["gamma"][-1, :, 0].copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_share_chains_and_agree_on_rhat(bnr):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        r, rg, rx, last = q.get(timeout=300)
        got[r] = (rg, rx, last)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    X, y = _data()
    one = bnr.Fit(X, y, 3, nburn=60, nsamples=40, num_chains=6, seed=17, x_transform=False, filename=None,
                  psrf_cutoff=1e9, return_state="gamma_xi")
    np.testing.assert_array_equal(got[0][0], got[1][0])                     # identical reduction on every rank
    np.testing.assert_allclose(got[0][0], one.rhatγ.γ, rtol=1e-12)         # == one process holding all 6 chains
    fin = np.isfinite(one.rhatξ.ξ)
    np.testing.assert_allclose(got[0][1][fin], np.asarray(one.rhatξ.ξ)[fin], rtol=1e-12)
    np.testing.assert_array_equal(got[0][2], one.state["gamma"][-1, :, 0])  # rank 0's chain 1 == global chain 1
