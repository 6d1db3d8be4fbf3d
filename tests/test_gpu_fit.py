"""bnr_fit on the GPU against the restated reference control loops (oracle/psrf_loops.py) driven by the engine's OWN
draws: the random numbers are keyed by (chain, sweep number), so one long traced run provides draw(c, s) for every sweep
any bookkeeping scheme can reach; the restated generate_samples! / generate_samples_dbl! then say which sweeps must sit
in which table rows, how many sweeps get generated and what R-hat the rows give -- data-dependent PSRF decisions
included."""
import numpy as np
import pytest

from oracle import psrf_loops as PL

pytestmark = pytest.mark.gpu


def _toy(seed, V=6, n=40):
    rng = np.random.default_rng(seed)
    q = V * (V + 1) // 2
    X = rng.normal(size=(n, q)) * (rng.random((n, q)) < 0.7)
    y = 1.0 + X[:, :5].sum(axis=1) + rng.normal(0, 1.5, size=n)
    return X, y


CASES = [
    dict(nburn=40, nsamples=30), dict(nburn=40, nsamples=30, purge_burn=10), dict(nburn=20, nsamples=30),
    dict(nburn=24, nsamples=8, purge_burn=4), dict(nburn=30, nsamples=21, purge_burn=5), dict(nburn=12, nsamples=50, purge_burn=4),
    dict(mingen=40, maxgen=200), dict(mingen=10, maxgen=45), dict(mingen=14, maxgen=60, purge_burn=3),
]


@pytest.mark.parametrize("kw", CASES)
@pytest.mark.parametrize("cutoff", [1.02, 1.3, 1e9])
def test_fit_matches_reference_loops_on_real_draws(bnr, kw, cutoff):
    X, y = _toy(5)
    C, R, seed, smax = 3, 3, 21, 420
    with bnr.Engine(X, y, R, num_chains=C, seed=seed, trace_rows=smax + 1, trace_full_chains=0) as eng:
        eng.init_state()
        eng.run(smax)
        G = np.stack([eng.get_trace(c, "gamma", 0, smax + 1)[:, :, 0] for c in range(C)])
        Xi = np.stack([eng.get_trace(c, "xi", 0, smax + 1)[:, :, 0] for c in range(C)])

    def draw(c, s):
        return Xi[c, s], G[c, s]

    fit_kw = dict(num_chains=C, seed=seed, x_transform=False, filename=None, psrf_cutoff=cutoff, return_state="gamma_xi")
    try:
        if "mingen" in kw:
            want = PL.generate_samples_dbl(draw, C, kw["mingen"], kw["maxgen"], cutoff, kw.get("purge_burn"))
        else:
            want = PL.generate_samples(draw, C, kw["nburn"], kw["nsamples"], kw["nburn"] + kw["nsamples"], cutoff,
                                       kw.get("purge_burn"))
    except IndexError:
        # the reference indexes past its own table here (BoundsError): the library refuses instead of dropping rows
        with pytest.raises(bnr.BnrError, match="BoundsError"):
            bnr.Fit(X, y, R, **fit_kw, **kw)
        return
    assert want["stale_reads"] == 0 and max(want["sweeps"]) <= smax
    res = bnr.Fit(X, y, R, **fit_kw, **kw)
    assert res.extra["tot_generated"] == want["tot_generated"]
    assert res.burn_in == want["burn_in"] and res.sampled == want["sampled"]
    assert res.extra["n_psrf"] == len(want["history"])
    rows = want["rows"][0]
    assert len(res.state) == len(rows)
    for r, s in enumerate(rows):
        if s is not None:
            np.testing.assert_array_equal(res.state["gamma"][r, :, 0], G[0, s], err_msg="row %d sweep %d" % (r, s))
            np.testing.assert_array_equal(res.state["xi"][r, :, 0], Xi[0, s])
    for got, ref in ((res.rhatγ.γ, want["rhat_gamma"]), (res.rhatξ.ξ, want["rhat_xi"])):
        got, ref = np.asarray(got), np.asarray(ref)
        assert (np.isnan(got) == np.isnan(ref)).all()
        ok = np.isfinite(ref)
        np.testing.assert_allclose(got[ok], ref[ok], rtol=1e-9)
        assert (np.isinf(got) == np.isinf(ref)).all()


def test_fit_refuses_what_the_reference_cannot_index(bnr):
    X, y = _toy(6)
    with pytest.raises(bnr.BnrError, match="BoundsError"):
        bnr.Fit(X, y, 3, nburn=40, nsamples=30, purge_burn=1, num_chains=2, seed=1, x_transform=False, filename=None)


def test_fit_ess_and_c_level_summary(bnr):
    """bnr_fit's optional ESS (streamed with the R-hat window) equals the trace-based ESS of the same draws, and the
    device Summary statistics equal the order statistics of the returned table."""
    X, y = _toy(7)
    C, R, seed, nburn, nsamp, L = 4, 3, 9, 60, 48, 15
    res = bnr.Fit(X, y, R, nburn=nburn, nsamples=nsamp, num_chains=C, seed=seed, x_transform=False, filename=None,
                  psrf_cutoff=1e9, ess_max_lag=L, return_state="gamma_xi")
    assert res.extra["rhat_streamed"] and res.extra["ess"] is not None
    with bnr.Engine(X, y, R, num_chains=C, seed=seed, trace_rows=nburn + nsamp, trace_full_chains=0) as eng:
        eng.init_state()
        eng.run(nburn + nsamp - 1)
        ex, eg = eng.ess(nburn, nsamp, L)
    np.testing.assert_allclose(res.extra["ess"]["gamma"], eg, rtol=1e-7)
    ok = np.isfinite(ex)
    np.testing.assert_allclose(res.extra["ess"]["xi"][ok], ex[ok], rtol=1e-7)
    g = res.state["gamma"][nburn:nburn + nsamp, :, 0]
    ds = res.extra["device_summary"]
    gs = np.sort(g, axis=0)
    np.testing.assert_array_equal(ds["lower"], gs[0])          # round(48 * 0.025) = 1 -> first order statistic
    np.testing.assert_array_equal(ds["upper"], gs[46])         # round(48 * 0.975) = 47 (ties-to-even of 46.8)
    np.testing.assert_allclose(ds["mean"], g.mean(axis=0), rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("exchange", ["nccl", "peer"])
def test_fit_over_two_devices_equals_one_device(bnr, exchange, monkeypatch):
    """Library-side multi-GPU: chains sharded over the GPUs of ONE process, split-half moments (and ESS statistics)
    all-gathered inside libbnr (ncclAllGather; peer copies when NCCL cannot be loaded) -- same R-hat and same chain 1 as
    one device holding all chains."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    if exchange == "peer":
        monkeypatch.setenv("BNR_EXCHANGE", "peer")
    X, y = _toy(8)
    kw = dict(nburn=60, nsamples=40, seed=17, x_transform=False, filename=None, psrf_cutoff=1e9, ess_max_lag=15,
              return_state="gamma_xi")
    two = bnr.Fit(X, y, 3, num_chains=3, n_devices=2, **kw)
    one = bnr.Fit(X, y, 3, num_chains=6, n_devices=1, **kw)
    assert two.extra["n_devices"] == 2 and two.extra["total_chains"] == 6
    if exchange == "nccl":
        assert two.extra["exchange"] in ("nccl", "peer-copy")
    np.testing.assert_allclose(two.rhatγ.γ, one.rhatγ.γ, rtol=1e-12)
    np.testing.assert_array_equal(two.state["gamma"], one.state["gamma"])
    np.testing.assert_allclose(two.extra["ess"]["gamma"], one.extra["ess"]["gamma"], rtol=1e-9)
    assert len(two.extra["status"]) == 6
    # the doubling scheme with block moments (only chain 1 keeps a trace: the second device records nothing)
    kd = dict(mingen=40, maxgen=120, seed=17, x_transform=False, filename=None, psrf_cutoff=0.0, return_state="gamma_xi")
    two = bnr.Fit(X, y, 3, num_chains=3, n_devices=2, **kd)
    one = bnr.Fit(X, y, 3, num_chains=6, n_devices=1, **kd)
    assert two.extra["tot_generated"] == one.extra["tot_generated"] == 120 and two.extra["rhat_streamed"]
    np.testing.assert_allclose(two.rhatγ.γ, one.rhatγ.γ, rtol=1e-12)
    np.testing.assert_array_equal(two.state["gamma"], one.state["gamma"])
