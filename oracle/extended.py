"""Extended-precision (x87 80-bit long double, eps = 1.1e-19) restatement of the gamma draw  --  TEST INFRASTRUCTURE ONLY.

north_star asks for 1e-10 relative agreement of every conditional's mean, factor and draw; quantities behind the n x n
(or q x q) solve cannot meet a fixed 1e-10 when the system is ill conditioned, because the reference (LU) and the engine
(Cholesky) are both only backward stable: their forward errors are bounded by c * cond * eps_double.  This module is the
arbiter SURVEY section 7 asks for: the same formulas as bnr_oracle.update_gamma / update_gamma_qform
(src/gibbs.jl:420-438) evaluated in long double, against which BOTH the float64 oracle and the CUDA engine are measured,
so that the tolerance rule max(1e-10, C * cond * eps) in tests/test_gpu_parity*.py rests on a measured C."""
import math

import numpy as np

LD = np.longdouble
EPS64 = float(np.finfo(np.float64).eps)


def chol_ld(A):
    """Lower Cholesky factor in long double (right-looking, vectorised per column)."""
    A = np.array(A, dtype=LD)
    n = A.shape[0]
    L = np.zeros((n, n), dtype=LD)
    for j in range(n):
        d = A[j, j]
        if not d > 0:
            raise np.linalg.LinAlgError("not positive definite")
        ljj = np.sqrt(d)
        L[j, j] = ljj
        if j + 1 < n:
            col = A[j + 1:, j] / ljj
            L[j + 1:, j] = col
            A[j + 1:, j + 1:] -= np.outer(col, col)
    return L


def solve_lower_ld(L, b):
    x = np.array(b, dtype=LD)
    n = L.shape[0]
    for j in range(n):
        x[j] = x[j] / L[j, j]
        if j + 1 < n:
            x[j + 1:] -= L[j + 1:, j] * x[j]
    return x


def solve_upper_from_lower_ld(L, b):
    """Solve L' x = b."""
    x = np.array(b, dtype=LD)
    n = L.shape[0]
    for j in range(n - 1, -1, -1):
        x[j] = x[j] / L[j, j]
        if j > 0:
            x[:j] -= L[j, :j] * x[j]
    return x


def W_of_ld(u, lam):
    u = np.asarray(u, dtype=LD)
    full = u.T @ (np.asarray(lam, dtype=LD)[:, None] * u)
    V = u.shape[1]
    return np.concatenate([full[k:, k] for k in range(V)])


def update_gamma_ld(X, y, tau2, u_new, lam_prev, S_prev, mu_prev, z1, z2):
    """bnr_oracle.update_gamma (src/gibbs.jl:420-438) in long double; the n x n system is solved by Cholesky."""
    X = np.asarray(X, dtype=LD); y = np.asarray(y, dtype=LD)
    S = np.asarray(S_prev, dtype=LD); z1 = np.asarray(z1, dtype=LD); z2 = np.asarray(z2, dtype=LD)
    tau2 = LD(tau2); mu = LD(mu_prev)
    n = X.shape[0]
    W = W_of_ld(u_new, lam_prev)
    tau = np.sqrt(tau2)
    d1 = np.sqrt(tau2 * S) * z1
    Xt = X / tau
    a1 = (y - X @ W - mu) / tau
    a3 = Xt @ d1 + z2
    G = (Xt * (tau2 * S)[None, :]) @ Xt.T + np.eye(n, dtype=LD)
    L = chol_ld(G)
    a4 = solve_upper_from_lower_ld(L, solve_lower_ld(L, a1 - a3))
    gamma = d1 + (tau2 * S) * (Xt.T @ a4) + W
    return dict(W=W, G=G, L=L, a4=a4, gamma=gamma)


def update_gamma_qform_ld(X, y, tau2, u_new, lam_prev, S_prev, mu_prev, z):
    """bnr_oracle.update_gamma_qform in long double."""
    X = np.asarray(X, dtype=LD); y = np.asarray(y, dtype=LD)
    S = np.asarray(S_prev, dtype=LD); z = np.asarray(z, dtype=LD)
    tau2 = LD(tau2); mu = LD(mu_prev)
    W = W_of_ld(u_new, lam_prev)
    P = (X.T @ X + np.diag(1 / S)) / tau2
    L = chol_ld(P)
    b = X.T @ ((y - mu - X @ W) / tau2)
    w = solve_lower_ld(L, b)
    beta = solve_upper_from_lower_ld(L, w + z)
    return dict(W=W, P=P, L=L, beta=beta, gamma=W + beta)


def rel_err(got, want):
    """Norm-wise relative error max|got - want| / max|want| (both converted to long double)."""
    want = np.asarray(want, dtype=LD)
    got = np.asarray(got, dtype=LD)
    return float(np.max(np.abs(got - want)) / np.max(np.abs(want)))
