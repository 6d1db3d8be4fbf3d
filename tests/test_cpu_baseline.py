"""The CPU timing port (oracle/cpu_baseline.py) draws from the same conditionals as the pinned oracle."""
import math

import numpy as np
from scipy import stats

from oracle import bnr_oracle as O
from oracle import cpu_baseline as B


def test_vectorised_gig_matches_distribution():
    rng = np.random.default_rng(5)
    for chi, psi in ((0.004, 1.0), (1.3, 0.9), (9.0, 4.0)):
        x = B.gig_half_vec(np.full(6000, chi), psi, rng)
        pit = stats.invgauss.cdf(1.0 / x, math.sqrt(psi / chi) / psi, scale=psi)
        assert stats.kstest(pit, "uniform").pvalue > 1e-3, (chi, psi)
    # mixed branches in one call keep their positions
    chi = np.array([0.004, 1.3, 9.0] * 2000)
    x = B.gig_half_vec(chi, 1.0, rng)
    for i, c in enumerate((0.004, 1.3, 9.0)):
        pit = stats.invgauss.cdf(1.0 / x[i::3], math.sqrt(1.0 / c) / 1.0, scale=1.0)
        assert stats.kstest(pit, "uniform").pvalue > 1e-3


def test_port_transitions_follow_oracle_conditionals(golden):
    """Plug consecutive states of the timing port into the pinned oracle's conditionals (same battery as
    tests/test_oracle_golden.py applies to the Julia goldens): tau2/theta PIT, whitened gamma and u residuals,
    GIG PIT, xi calibration."""
    X, y = golden["test1.X"], golden["test1.y"]
    V, R = 19, 5
    port = B.ReferencePort(X, y, R, seed=3)
    for _ in range(30):
        port.sweep()
    p_tau, p_th, res_g, res_u, pit_s = [], [], [], [], []
    psum = xsum = pvar = 0.0
    for it in range(120):
        a = dict(port.st)
        b = port.sweep()
        t = O.update_tau2(X, y, V, a["mu"], a["gamma"], a["u"], a["lam"], a["S"], 1.0)
        p_tau.append(stats.gamma.cdf(t["scale"] / b["tau2"], t["shape"]))
        th = O.update_theta(b["S"], 1.0, 1.0, V, 1.0)
        p_th.append(stats.gamma.cdf(b["theta"] / th["scale"], th["shape"]))
        W = O.W_of(b["u"], a["lam"])
        if it % 4 == 0:
            m, P = O.gamma_conditional_moments(X, y, b["tau2"], W, a["S"], a["mu"])
            res_g.append(np.linalg.cholesky(P).T @ (b["gamma"] - m))
        chi = (b["gamma"] - W) ** 2 / b["tau2"]
        pit_s.append(stats.invgauss.cdf(1.0 / b["S"], np.sqrt(a["theta"] / chi) / a["theta"], scale=a["theta"]))
        for k in range(0, V, 3):
            r = O.update_u_xi_node(k, V, b["tau2"], a["u"], a["lam"], a["S"], a["gamma"], a["Delta"], a["M"],
                                   0.5, np.zeros(R), literal=False)
            p1 = 1.0 - r["w"]
            psum += p1; pvar += p1 * (1 - p1); xsum += b["xi"][k]
            if b["xi"][k] == 1:
                res_u.append(r["chol"].T @ (b["u"][:, k] - r["mu_t"]))
    assert stats.kstest(p_tau, "uniform").pvalue > 1e-3
    assert stats.kstest(p_th, "uniform").pvalue > 1e-3
    assert stats.kstest(np.concatenate(res_g), "norm").pvalue > 1e-3
    assert stats.kstest(np.concatenate(res_u), "norm").pvalue > 1e-3
    assert stats.kstest(np.concatenate(pit_s), "uniform").pvalue > 1e-3
    assert abs((xsum - psum) / math.sqrt(max(pvar, 1e-9))) < 4.0


def test_index_helper():
    il, ik = B.lower_triangle_idx(5)
    M = np.arange(25.0).reshape(5, 5)
    np.testing.assert_array_equal(M[il, ik], O.lower_triangle(M))
