"""The N>1 path on CPU: two gloo ranks gather per-rank split-half moments exactly the way bench.py / a multi-GPU
Fit does with NCCL, and every rank reduces them to the same R-hat as a single-process computation over all
chains.  (The device-side reduction kernel itself is covered by the -m gpu tests; here the exchange layout --
[chain][half][param][mean, M2], ranks concatenated in global-chain order -- is what is pinned.)"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import bnr_oracle as O


def _moments(traces):
    """traces: (draws, params, chains) -> [chain][half][param][2] like bnr_moments_device."""
    n, p, c = traces.shape
    h = n // 2
    out = np.empty((c, 2, p, 2))
    for ch in range(c):
        for half, sl in enumerate((slice(0, h), slice(n - h, n))):
            x = traces[sl, :, ch]
            m = x.mean(axis=0)
            out[ch, half, :, 0] = m
            out[ch, half, :, 1] = ((x - m) ** 2).sum(axis=0)
    return out


def _worker(rank, world, port, traces, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    per = traces.shape[2] // world
    mine = torch.from_numpy(_moments(traces[:, :, rank * per:(rank + 1) * per]).ravel().copy())
    allm = torch.empty(mine.numel() * world, dtype=torch.float64)
    dist.all_gather_into_tensor(allm, mine)
    mom = allm.numpy().reshape(world * per, 2, traces.shape[1], 2)
    h = traces.shape[0] // 2
    mean = mom[:, :, :, 0].reshape(-1, traces.shape[1])
    m2 = mom[:, :, :, 1].reshape(-1, traces.shape[1])
    # ESS statistics: per-rank autocovariance sums [L+1][param] and chain means [chain][param], gathered the same way
    L = 15
    loc = traces[:, :, rank * per:(rank + 1) * per]
    st = [O.ess_stats(loc[:, p, :], L) for p in range(traces.shape[1])]
    a_mine = torch.from_numpy(np.stack([s_[0] for s_ in st], axis=1).ravel().copy())           # [L+1][param]
    m_mine = torch.from_numpy(np.stack([s_[1] for s_ in st], axis=1).ravel().copy())           # [chain][param]
    a_all = torch.empty(a_mine.numel() * world, dtype=torch.float64)
    m_all = torch.empty(m_mine.numel() * world, dtype=torch.float64)
    dist.all_gather_into_tensor(a_all, a_mine)
    dist.all_gather_into_tensor(m_all, m_mine)
    A = a_all.numpy().reshape(world, L + 1, traces.shape[1])
    M = m_all.numpy().reshape(world * per, traces.shape[1])
    ess = np.array([O.ess_from_stats(A[:, :, p], M[:, p], traces.shape[0], L) for p in range(traces.shape[1])])
    q.put((rank, (O.rhat_from_moments(mean, m2, h), ess)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_moment_gather_gives_global_rhat():
    rng = np.random.default_rng(0)
    draws, params, chains = 101, 7, 6
    traces = rng.normal(size=(draws, params, chains)).cumsum(axis=0) * 0.1 + rng.normal(size=(1, params, chains))
    want = O.rhat(traces)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, traces, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want_ess = np.array([O.ess_geyer(traces[:, p, :], 15) for p in range(params)])
    for r in range(2):
        np.testing.assert_allclose(got[r][0], want, rtol=1e-11)
        np.testing.assert_allclose(got[r][1], want_ess, rtol=1e-10)
    np.testing.assert_array_equal(got[0][0], got[1][0])      # every rank gets the identical reduction
    np.testing.assert_array_equal(got[0][1], got[1][1])
