"""The float64 oracle's gamma draw against the long-double arbiter (oracle/extended.py): its forward error follows
c * cond * eps with a small c -- the basis of the tolerance rule used by the GPU parity tests."""
import numpy as np
import pytest

from oracle import bnr_oracle as O
from oracle import extended as E


def _problem(seed, V, R, n, s_scale):
    rng = np.random.default_rng(seed)
    q = V * (V + 1) // 2
    X = rng.normal(size=(n, q))
    y = 3.0 + X[:, :5].sum(axis=1) + rng.normal(size=n)
    u = rng.normal(size=(R, V))
    lam = rng.choice([0.0, 1.0, -1.0], size=R)
    S = (rng.gamma(1.0, size=q) + 1e-3) * s_scale
    return X, y, u, lam, S, rng.normal(size=q), rng.normal(size=n)


@pytest.mark.skipif(np.finfo(np.longdouble).eps > 1e-18, reason="no 80-bit long double on this platform")
@pytest.mark.parametrize("s_scale", [1.0, 1e3, 1e6])
def test_float64_oracle_error_is_cond_times_eps(s_scale):
    X, y, u, lam, S, z1, z2 = _problem(1, 10, 3, 60, s_scale)
    tau2, mu = 1.7, 0.3
    ld = E.update_gamma_ld(X, y, tau2, u, lam, S, mu, z1, z2)
    f64 = O.update_gamma(X, y, tau2, u, lam, S, mu, z1, z2)
    cond = np.linalg.cond(f64["G"])
    err = E.rel_err(f64["gamma"], ld["gamma"])
    assert err <= max(1e-13, 8.0 * cond * E.EPS64), (err, cond)
    ldq = E.update_gamma_qform_ld(X, y, tau2, u, lam, S, mu, z1)
    f64q = O.update_gamma_qform(X, y, tau2, u, lam, S, mu, z1)
    condq = np.linalg.cond(f64q["P"])
    errq = E.rel_err(f64q["gamma"], ldq["gamma"])
    assert errq <= max(1e-13, 8.0 * condq * E.EPS64), (errq, condq)


def test_long_double_solvers():
    rng = np.random.default_rng(0)
    A = rng.normal(size=(12, 12))
    A = A @ A.T + 12 * np.eye(12)
    L = E.chol_ld(A)
    assert float(np.max(np.abs(L @ L.T - A))) < 1e-15 * 12 * 30
    b = rng.normal(size=12)
    x = E.solve_upper_from_lower_ld(L, E.solve_lower_ld(L, b))
    np.testing.assert_allclose(np.asarray(A @ np.asarray(x, dtype=np.float64)), b, rtol=1e-12, atol=1e-12)
