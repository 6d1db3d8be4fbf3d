"""Pin the CPU oracle against the reference's own golden vectors.

Exact pins: R-hat (src/convergence.jl:4-65) and Summary (src/gibbs.jl:1214-1250) reproduce the stored
`res/res2.rhat*` and `out/out2` tables.  Statistical pins: the 399 stored consecutive transitions of
`res2` (fit on test/data/test1.csv) are plugged into every conditional of oracle/bnr_oracle.py; a wrong
row index (i vs i-1), column order, GIG parameter role or q-convention makes these fail loudly.
"""
import math

import numpy as np
import pytest
from scipy import stats

from oracle import bnr_oracle as O

P_MIN = 1e-3   # deterministic data: a fixed, conservative rejection level


def _row(g, key, i):
    st = dict(tau2=g[key + ".tau2"][i, 0, 0], u=g[key + ".u"][i], xi=g[key + ".xi"][i, :, 0],
              gamma=g[key + ".gamma"][i, :, 0], S=g[key + ".S"][i, :, 0], theta=g[key + ".theta"][i, 0, 0],
              Delta=g[key + ".Delta"][i, 0, 0], M=g[key + ".M"][i], mu=g[key + ".mu"][i, 0, 0],
              lam=g[key + ".lambda"][i, :, 0], pi=g[key + ".pi"][i])
    return st


@pytest.mark.parametrize("key,lo", [("res", 300), ("res2", 200)])
def test_rhat_exact(golden, key, lo):
    n = int(golden[key + ".sampled"])
    assert lo == int(golden[key + ".burn_in"])
    for var, name in (("gamma", "rhat_gamma"), ("xi", "rhat_xi")):
        tr = golden[key + "." + var][lo:lo + n, :, 0][:, :, None]
        got = O.rhat(tr)
        want = golden[key + "." + name]
        np.testing.assert_allclose(got, want, rtol=1e-12, equal_nan=True)
        # and through streaming moments (what the device accumulates)
        h = n // 2
        halves = np.stack([tr[:h, :, 0], tr[n - h:, :, 0]])
        mean = halves.mean(axis=1)
        m2 = ((halves - mean[:, None]) ** 2).sum(axis=1)
        got2 = O.rhat_from_moments(mean, m2, h)
        ok = np.isfinite(want)
        np.testing.assert_allclose(got2[ok], want[ok], rtol=1e-11)


@pytest.mark.parametrize("key,okey", [("res", "out"), ("res2", "out2")])
def test_summary_exact(golden, key, okey):
    lo, n = int(golden[key + ".burn_in"]), int(golden[key + ".sampled"])
    s = O.summary(golden[key + ".gamma"][lo:lo + n, :, 0], golden[key + ".xi"][lo:lo + n, :, 0])
    for f in ("node1", "node2", "estimate", "lower_bound", "upper_bound", "probability"):
        np.testing.assert_array_equal(s[f], golden[okey + "." + f], err_msg=f)
    assert s["ci_level"] == int(golden[okey + ".ci_level"])


def test_index_maps():
    V = 6
    A = np.arange(V * V, dtype=float).reshape(V, V)
    A = A + A.T
    v = O.lower_triangle(A)
    assert v.shape[0] == V * (V + 1) // 2
    L = O.create_lower_tri(v, V)
    np.testing.assert_array_equal(L, np.tril(A))
    for k in range(V):
        for l in range(k, V):
            assert v[O.tri_index(l, k, V)] == A[l, k]
    # node_edge_indices returns the V-1 off-diagonal entries of node k in reference order
    for k in range(V):
        idx = O.node_edge_indices(k, V)
        want = [A[k, l] for l in range(V) if l != k]
        np.testing.assert_array_equal(v[idx], want)


def test_initial_state_constants(golden):
    """st1 row 1 (src/gibbs.jl:199-217): theta=0.5, Delta=0.5, mu=1, tau2=1 and the stored draws have
    the supports the oracle assumes."""
    g = golden
    assert g["st1.theta"][0, 0, 0] == 0.5 and g["st1.Delta"][0, 0, 0] == 0.5
    assert g["st1.mu"][0, 0, 0] == 1.0 and g["st1.tau2"][0, 0, 0] == 1.0
    assert set(np.unique(g["st1.lambda"][0, :, 0])) <= {-1.0, 0.0, 1.0}
    assert set(np.unique(g["st1.xi"][0, :, 0])) <= {0.0, 1.0}
    np.testing.assert_allclose(g["st1.pi"][0].sum(axis=1), 1.0, rtol=1e-12)
    np.testing.assert_allclose(g["st2.tau2"][1, 0, 0], 37.606403334971134)


@pytest.fixture(scope="module")
def transitions(golden):
    g = golden
    X, y = g["test1.X"], g["test1.y"]
    T = g["res2.gamma"].shape[0]
    return g, X, y, T


def test_pit_tau2_theta_Delta_mu(transitions):
    g, X, y, T = transitions
    V, R = 19, 5
    p_tau, p_th, p_de, z_mu = [], [], [], []
    for i in range(1, T):
        a, b = _row(g, "res2", i - 1), _row(g, "res2", i)
        t = O.update_tau2(X, y, V, a["mu"], a["gamma"], a["u"], a["lam"], a["S"], 1.0)
        p_tau.append(stats.gamma.cdf(t["scale"] / b["tau2"], t["shape"]))
        th = O.update_theta(b["S"], 1.0, 1.0, V, 1.0)
        p_th.append(stats.gamma.cdf(b["theta"] / th["scale"], th["shape"]))
        de = O.update_Delta(b["xi"], 1.0, 1.0, 1.0, 1.0)
        p_de.append(stats.beta.cdf(b["Delta"], de["a"], de["b"]))
        mu = O.update_mu(X @ b["gamma"], y, b["tau2"], 0.0)
        z_mu.append((b["mu"] - mu["mean"]) / mu["sd"])
    assert stats.kstest(p_tau, "uniform").pvalue > P_MIN
    assert stats.kstest(p_th, "uniform").pvalue > P_MIN
    assert stats.kstest(p_de, "uniform").pvalue > P_MIN
    assert stats.kstest(z_mu, "norm").pvalue > P_MIN
    assert abs(np.std(z_mu) - 1) < 0.12


def test_whitened_gamma(transitions):
    g, X, y, T = transitions
    res = []
    for i in range(1, T, 4):   # every 4th transition keeps the CPU suite fast (q=190 solves)
        a, b = _row(g, "res2", i - 1), _row(g, "res2", i)
        W = O.W_of(b["u"], a["lam"])
        m, P = O.gamma_conditional_moments(X, y, b["tau2"], W, a["S"], a["mu"])
        L = np.linalg.cholesky(P)
        res.append(L.T @ (b["gamma"] - m))
    res = np.concatenate(res)
    assert abs(res.mean()) < 0.02 and abs(res.std() - 1) < 0.02
    assert stats.kstest(res, "norm").pvalue > P_MIN


def test_u_xi_conditional(transitions):
    g, X, y, T = transitions
    V, R = 19, 5
    res, psum, xsum, pvar = [], 0.0, 0.0, 0.0
    for i in range(1, T, 3):
        a, b = _row(g, "res2", i - 1), _row(g, "res2", i)
        for k in range(V):
            r = O.update_u_xi_node(k, V, b["tau2"], a["u"], a["lam"], a["S"], a["gamma"], a["Delta"], a["M"],
                                   0.5, np.zeros(R), literal=False)
            p1 = 1.0 - r["w"]
            psum += p1
            pvar += p1 * (1 - p1)
            xsum += b["xi"][k]
            if b["xi"][k] == 1:
                res.append(r["chol"].T @ (b["u"][:, k] - r["mu_t"]))
            else:
                assert np.all(b["u"][:, k] == 0)
    res = np.concatenate(res)
    assert abs((xsum - psum) / math.sqrt(pvar)) < 4.0
    assert abs(res.std() - 1) < 0.03 and stats.kstest(res, "norm").pvalue > P_MIN


def test_xi_logodds_literal_equals_woodbury(transitions):
    g, X, y, T = transitions
    V, R = 19, 5
    for i in (1, 57, 203, 398):
        a, b = _row(g, "res2", i - 1), _row(g, "res2", i)
        for k in (0, 7, 18):
            args = (k, V, b["tau2"], a["u"], a["lam"], a["S"], a["gamma"], a["Delta"], a["M"], 0.3, np.ones(R))
            lit = O.update_u_xi_node(*args, literal=True)
            fast = O.update_u_xi_node(*args, literal=False)
            assert abs(lit["log_odds"] - fast["log_odds"]) < 1e-8 * max(1.0, abs(lit["log_odds"]))
            np.testing.assert_allclose(lit["u"], fast["u"], rtol=1e-12, atol=1e-14)


def test_gig_pit_and_branches(transitions):
    g, X, y, T = transitions
    pit = []
    for i in range(1, T, 2):
        a, b = _row(g, "res2", i - 1), _row(g, "res2", i)
        W = O.W_of(b["u"], a["lam"])
        chi = (b["gamma"] - W) ** 2 / b["tau2"]
        psi = a["theta"]
        # 1/S ~ InverseGaussian(mean sqrt(psi/chi), shape psi)
        mean = np.sqrt(psi / chi)
        pit.append(stats.invgauss.cdf(1.0 / b["S"], mean / psi, scale=psi))
    pit = np.concatenate(pit)
    assert stats.kstest(pit, "uniform").pvalue > P_MIN


def test_gig_sampler_distribution():
    """The restated Hormann-Leydold sampler (src/gig.jl) draws GIG(1/2, chi, psi) in all three branches."""
    rng = np.random.default_rng(7)
    for chi, psi, want in ((0.004, 1.0, "concave"), (1.3, 0.9, "noshift"), (9.0, 4.0, "shift")):
        xs = []
        for _ in range(4000):
            x, br = O.sample_gig(0.5, chi, psi, O.UniformStream(rng.random(400)))
            assert br == want
            xs.append(x)
        pit = stats.invgauss.cdf(1.0 / np.asarray(xs), math.sqrt(psi / chi) / psi, scale=psi)
        assert stats.kstest(pit, "uniform").pvalue > P_MIN, want


def test_lambda_pi_M(transitions):
    g, X, y, T = transitions
    V, R = 19, 5
    exp_counts = np.zeros(3)
    obs = np.zeros(3)
    pit_pi, tr = [], []
    for i in range(1, T, 2):
        a, b = _row(g, "res2", i - 1), _row(g, "res2", i)
        la = O.update_lambda(b["gamma"], b["u"], b["S"], b["tau2"], a["lam"], a["pi"], np.full(R, 0.5))
        w = la["weights"] / la["weights"].sum(axis=1, keepdims=True)
        exp_counts += w.sum(axis=0)
        for r in range(R):
            obs[O.LAMBDA_VALUES.index(b["lam"][r])] += 1
        pi = O.update_pi(b["lam"], 1.01, np.ones((R, 3)))
        al = pi["alpha"]
        pit_pi.extend(stats.beta.cdf(b["pi"][:, 0], al[:, 0], al[:, 1] + al[:, 2]))
        m = O.update_M(b["u"], b["xi"], 10, np.ones(R), np.zeros(R * (R - 1) // 2))
        tr.append(np.trace(m["Psi"] @ np.linalg.inv(b["M"])) / (m["df"] * R))
    chi2 = ((obs - exp_counts) ** 2 / exp_counts).sum()
    assert chi2 < 20.0, (obs, exp_counts)
    assert stats.kstest(pit_pi, "uniform").pvalue > P_MIN
    assert abs(np.mean(tr) - 1.0) < 0.03


def test_sweep_from_injected_draws_is_self_consistent():
    """gibbs_sweep/initialize_state run end to end and the literal and Woodbury xi-odds agree."""
    rng = np.random.default_rng(3)
    V, R, n, K = 5, 3, 12, 64
    q = V * (V + 1) // 2
    X = rng.normal(size=(n, q))
    y = rng.normal(size=n) + 3
    il = O.init_layout(V, R)
    inj0 = np.abs(rng.normal(size=il["_total"])) + 0.05
    for name in ("lambda", "xi"):
        o, s = il[name]
        inj0[o:o + s] = rng.random(s)
    for name in ("u", "gamma"):
        o, s = il[name]
        inj0[o:o + s] = rng.normal(size=s)
    st = O.initialize_state(V, R, O.DEFAULT_HYPER, inj0)
    lay = O.draw_layout(n, V, R, K)
    for it in range(3):
        inj = rng.normal(size=lay["_total"])
        for name in ("tau2", "theta", "pi", "Delta"):
            o, s = lay[name]
            inj[o:o + s] = rng.gamma(3.0, size=s)
        o, s = lay["M"]
        inj[o:o + R] = rng.chisquare(10, size=R)
        for name in ("S", "lambda"):
            o, s = lay[name]
            inj[o:o + s] = rng.random(s)
        o, s = lay["uxi"]
        inj[o:o + s:R + 1] = rng.random(V)
        a, _ = O.gibbs_sweep(st, X, y, V, R, O.DEFAULT_HYPER, inj, K, literal=True)
        b, _ = O.gibbs_sweep(st, X, y, V, R, O.DEFAULT_HYPER, inj, K, literal=False)
        for key in a:
            np.testing.assert_allclose(a[key], b[key], rtol=1e-9, atol=1e-12, err_msg=key)
        st = a
    assert np.all(st["S"] > 0) and st["tau2"] > 0


def test_qform_gamma_draw_has_the_reference_conditional_law():
    """The q x q precision draw (oracle.update_gamma_qform, BASELINE's formulation) and the reference's n x n
    Bhattacharya draw (update_gamma, src/gibbs.jl:420-438) are the same Gaussian: equal mean (all normals 0)
    and equal covariance (the n-form is linear in (z1, z2): gamma = mean + A1 z1 + A2 z2, A1 A1' + A2 A2' = P^-1)."""
    rng = np.random.default_rng(5)
    V, R, n = 5, 3, 9
    q = V * (V + 1) // 2
    X = rng.normal(size=(n, q))
    y = rng.normal(size=n) + X[:, 0]
    u = rng.normal(size=(R, V)); lam = np.array([1.0, 0.0, -1.0]); S = rng.gamma(1.0, size=q) + 0.05
    tau2, mu = 1.7, 0.3
    zq, zn = np.zeros(q), np.zeros(n)
    a = O.update_gamma(X, y, tau2, u, lam, S, mu, zq, zn)
    b = O.update_gamma_qform(X, y, tau2, u, lam, S, mu, zq)
    np.testing.assert_allclose(a["gamma"], b["gamma"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(b["mean"], b["gamma"], rtol=0, atol=0)
    A1 = np.stack([O.update_gamma(X, y, tau2, u, lam, S, mu, np.eye(q)[j], zn)["gamma"] - a["gamma"] for j in range(q)], axis=1)
    A2 = np.stack([O.update_gamma(X, y, tau2, u, lam, S, mu, zq, np.eye(n)[i])["gamma"] - a["gamma"] for i in range(n)], axis=1)
    cov_n = A1 @ A1.T + A2 @ A2.T
    cov_q = np.linalg.inv(b["P"])
    np.testing.assert_allclose(cov_n, cov_q, rtol=1e-8, atol=1e-10)
    # and the q-form draw itself is mean + L^-T z
    z = rng.normal(size=q)
    c = O.update_gamma_qform(X, y, tau2, u, lam, S, mu, z)
    np.testing.assert_allclose(c["gamma"], b["gamma"] + np.linalg.solve(b["L"].T, z), rtol=1e-10, atol=1e-12)
    m, P = O.gamma_conditional_moments(X, y, tau2, b["W"], S, mu)
    np.testing.assert_allclose(m, b["mean"], rtol=1e-9, atol=1e-11)


def test_ess_estimator_known_answers():
    """oracle.ess_geyer: iid draws give ESS ~ n m; an AR(1) chain with coefficient phi gives n m (1-phi)/(1+phi);
    the direct-lag form equals the FFT form used elsewhere."""
    rng = np.random.default_rng(0)
    n, m = 4000, 4
    iid = rng.normal(size=(n, m))
    e = O.ess_geyer(iid)
    assert 0.85 * n * m < e < 1.2 * n * m
    phi = 0.7
    x = np.zeros((n, m))
    eps = rng.normal(size=(n, m))
    for t in range(1, n):
        x[t] = phi * x[t - 1] + eps[t]
    e = O.ess_geyer(x, max_lag=255)
    want = n * m * (1 - phi) / (1 + phi)
    assert 0.8 * want < e < 1.25 * want
    # FFT cross-check of the autocovariances behind it
    xc = x - x.mean(axis=0)
    f = np.fft.rfft(xc, n=2 * n, axis=0)
    ac_fft = np.fft.irfft(f * np.conj(f), axis=0)[:8] / n
    ac_dir = np.stack([(xc[: n - t] * xc[t:]).sum(axis=0) / n for t in range(8)])
    np.testing.assert_allclose(ac_fft, ac_dir, rtol=1e-9, atol=1e-12)
    # a stuck chain (zero variance) is NaN, never a crash
    assert math.isnan(O.ess_geyer(np.ones((100, 2))))
