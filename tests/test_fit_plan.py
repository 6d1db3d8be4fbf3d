"""The control flow of bnr_fit (csrc/bnr_fit.cu: run! with the purge ring, the "extend burn-in" and "doubling" PSRF
loops) checked WITHOUT a GPU against the restated reference loops (oracle/psrf_loops.py): bnr_fit_plan returns the
engine operations a fit issues for given PSRF outcomes; replaying them on a table simulator must put the same sweeps in
the same rows, hand the same rows to every R-hat evaluation and generate the same number of sweeps as the reference."""
import ctypes as C
import itertools

import numpy as np
import pytest

from oracle import psrf_loops as PL


def plan(bnr, psrf_max, **kw):
    from bnr_b200 import capi
    L = bnr.lib()
    p = capi.FitParams()
    L.bnr_fit_default_params(C.byref(p))
    p.base.n, p.base.V, p.base.R, p.base.num_chains = 10, 4, 3, 2
    for k, v in kw.items():
        setattr(p, k, v)
    pm = np.asarray(psrf_max, dtype=np.float64)
    n_ops = C.c_int64()
    info = capi.FitInfo()
    cap = 1 << 16
    ops = np.zeros((cap, 4), dtype=np.int64)
    capi.check_fit(L.bnr_fit_plan(C.byref(p), pm.ctypes.data_as(C.POINTER(C.c_double)), len(pm),
                                  ops.ctypes.data_as(C.POINTER(C.c_int64)), cap, C.byref(n_ops), C.byref(info)))
    return ops[:n_ops.value], info


class Sim:
    """What the engine does with the operations (row / sweep bookkeeping of bnr_run, bnr_set_trace_row,
    bnr_copy_trace_rows, the moment windows)."""

    def __init__(self):
        self.table, self.trace_row, self.sweep = [], 0, 0
        self.win = self.blk = None
        self.psrf_used, self.gx_all = [], None

    def replay(self, ops):
        for op, a, b, c in ops.tolist():
            if op == 0:
                self.table, self.gx_all = [None] * a, bool(b)
            elif op == 1:
                self.table[0], self.trace_row, self.sweep = 0, 1, 0
            elif op == 2:
                for _ in range(a):
                    self.sweep += 1
                    if self.trace_row < len(self.table):
                        self.table[self.trace_row] = self.sweep
                    self.trace_row += 1
            elif op == 3:
                self.trace_row = a
            elif op == 4:
                self.table[a:a + c] = list(self.table[b:b + c])
            elif op == 5:
                self.win = (a, b)
            elif op == 6:
                self.blk = (a, b, c)
            elif op == 7:
                assert self.gx_all, "trace-based R-hat needs every chain's trace"
                self.psrf_used.append(list(self.table[a:a + b]))
            elif op == 8:
                f, n = self.win
                assert f + n - 1 == self.sweep, "the streamed window must end at the last sweep run"
                self.psrf_used.append(list(range(f, f + n)))
            elif op == 9:
                f, bl, cnt = self.blk
                assert a + b <= cnt and f + (a + b) * bl - 1 == self.sweep
                self.psrf_used.append(list(range(f + a * bl, f + (a + b) * bl)))
        return self


def _feed(values):
    it = iter(values)
    log = []

    def rhat_fn(used):
        assert all(u == used[0] for u in used)          # every chain holds the same sweeps in the same rows
        log.append(list(used[0]))
        v = next(it, 0.0)
        return np.array([v]), np.array([v])
    return rhat_fn, log


TRAD = [(nburn, nsamp, purge) for nburn, nsamp in ((40, 30), (30, 30), (12, 50), (100, 20), (24, 8), (7, 5), (12, 5))
        for purge in (None, 1, 3, 4, 5, 8, 10, 25)]


@pytest.mark.parametrize("nburn,nsamp,purge", TRAD)
@pytest.mark.parametrize("outcomes", [(0.0,), (9.0, 0.0), (9.0, 9.0, 9.0), (float("nan"),)])
def test_traditional_loop_matches_reference(bnr, nburn, nsamp, purge, outcomes):
    from bnr_b200 import capi
    rhat_fn, want_used = _feed(outcomes)
    kw = dict(nburn=nburn, nsamples=nsamp, psrf_cutoff=1.01, purge_burn=purge or 0)
    try:
        want = PL.generate_samples(None, 2, nburn, nsamp, nburn + nsamp, 1.01, purge, rhat_fn=rhat_fn)
    except IndexError:
        # the reference runs past its own table (BoundsError): the library must refuse, not silently drop rows
        with pytest.raises(capi.BnrError, match="BoundsError"):
            plan(bnr, outcomes, **kw)
        return
    ops, info = plan(bnr, outcomes, **kw)
    if want["stale_reads"]:
        pytest.skip("the reference itself continues from a stale row for this combination (documented divergence)")
    sim = Sim().replay(ops)
    assert info.tot_generated == want["tot_generated"] and info.burn_in == want["burn_in"]
    assert info.sampled == want["sampled"] and info.rows == len(want["rows"][0])
    assert sim.sweep == want["sweeps"][0]
    assert sim.psrf_used == want_used
    if info.streamed == 0:
        assert sim.table[:info.rows] == want["rows"][0]
    else:
        # chain 1 keeps its trace in every mode: same rows
        assert sim.table[:info.rows] == want["rows"][0]


@pytest.mark.parametrize("mingen,maxgen", [(40, 200), (40, 120), (20, 20), (10, 35), (14, 60), (8, 64), (9, 50)])
@pytest.mark.parametrize("purge", [None, 3, 5])
@pytest.mark.parametrize("outcomes", [(0.0,), (9.0, 0.0), (9.0, float("nan"), 9.0, 0.5), (9.0,) * 12])
def test_doubling_loop_matches_reference(bnr, mingen, maxgen, purge, outcomes):
    from bnr_b200 import capi
    rhat_fn, want_used = _feed(outcomes)
    kw = dict(mingen=mingen, maxgen=maxgen, psrf_cutoff=1.01, purge_burn=purge or 0)
    try:
        want = PL.generate_samples_dbl(None, 2, mingen, maxgen, 1.01, purge, rhat_fn=rhat_fn)
    except IndexError:
        with pytest.raises(capi.BnrError, match="BoundsError"):
            plan(bnr, outcomes, **kw)
        return
    ops, info = plan(bnr, outcomes, **kw)
    if want["stale_reads"]:
        pytest.skip("the reference itself continues from a stale row for this combination (documented divergence)")
    sim = Sim().replay(ops)
    assert info.tot_generated == want["tot_generated"] and info.burn_in == want["burn_in"]
    assert info.sampled == want["sampled"] and info.rows == len(want["rows"][0])
    assert sim.sweep == want["sweeps"][0]
    assert sim.psrf_used == want_used
    assert sim.table[:info.rows] == want["rows"][0]


def test_plan_rejects_bad_arguments(bnr):
    from bnr_b200 import capi
    with pytest.raises(capi.BnrError):
        plan(bnr, (0.0,), nburn=10, nsamples=0)
    with pytest.raises(capi.BnrError):
        plan(bnr, (0.0,), nburn=10, nsamples=5, purge_burn=-1)
