"""CPU oracle for the Gibbs hot path of BayesianNetworkRegression.jl  --  TEST INFRASTRUCTURE ONLY.

This file is a NumPy/FP64 restatement of the reference's algorithm (never of its code) and exists
solely so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
can check or time against it.  Nothing in the product package may import it; the product path is
the CUDA library (libbnr.so) and fails loudly when that is missing.

Parity status
-------------
* Julia is not installed in the build image, so the reference itself cannot be executed here.
* Pinned EXACTLY by the reference's own goldens (test/data/gen_test_results.jld2, decoded into
  tests/golden/golden.npz by tests/golden/make_golden.py): `rhat`, `summary`.
* Pinned STATISTICALLY by the same goldens (399 stored consecutive transitions of `res2`):
  every conditional below (PIT / whitened-residual battery in tests/test_oracle_golden.py).
* Draw-level equality with Julia's Xoshiro stream and Distributions.jl's variate algorithms
  (Gamma, Beta, Dirichlet, InverseWishart, categorical): PARITY UNPINNED -- those algorithms live
  in un-vendored third-party packages (Distributions 0.25.66, StatsBase 0.33.21, Project.toml:29-44,
  no Manifest).  The oracle therefore takes the *basic variates* (standard normals, uniforms,
  unit-scale gamma variates) as injected inputs; conditional parameters, means, Cholesky factors and
  the draws as functions of those variates are what is compared at 1e-10.

Conventions: u has shape (R, V) (reference state.u[i,:,:]); gamma/S have length q = V(V+1)/2 in the
order of src/utils.jl:40-57 (column k = 1..V, rows l = k..V, diagonal included); pi columns are
P(lambda=0), P(+1), P(-1) (src/gibbs.jl:161-165,207,610).
"""
import math
import numpy as np

EPS10 = 10.0 * np.finfo(np.float64).eps


# ------------------------------------------------------------------------------------------------
# index maps  (src/utils.jl:17-27, 40-57)
# ------------------------------------------------------------------------------------------------
def tri_index(l, k, V):
    """0-based position in the q-vector of matrix entry (l, k), l >= k (0-based)."""
    return k * V - (k * (k - 1)) // 2 + (l - k)


def lower_triangle(mat):
    """src/utils.jl:40-57: column-major lower triangle INCLUDING the diagonal."""
    V = mat.shape[0]
    return np.concatenate([mat[k:, k] for k in range(V)])


def create_lower_tri(vec, V):
    """src/utils.jl:17-27: inverse of lower_triangle (strict upper part left at zero)."""
    mat = np.zeros((V, V), dtype=np.asarray(vec).dtype)
    i = 0
    for k in range(V):
        mat[k:, k] = vec[i:i + V - k]
        i += V - k
    return mat


def setup_X(X_list):
    """src/gibbs.jl:239-247 with x_transform=true: one row per sample = lower_triangle(X[i])."""
    return np.stack([lower_triangle(np.asarray(x, dtype=np.float64)) for x in X_list])


def W_of(u, lam):
    """lower_triangle(u' diag(lam) u)  (src/gibbs.jl:219-221, 271, 421, 455)."""
    return lower_triangle(u.T @ (lam[:, None] * u))


def node_edge_indices(k, V):
    """Positions in the q-vector of the V-1 off-diagonal entries of node k, in the order the
    reference builds gamma_k / H (src/gibbs.jl:300-309): (k,0..k-1) then (k+1..V-1,k)."""
    idx = [tri_index(k, l, V) for l in range(k)] + [tri_index(l, k, V) for l in range(k + 1, V)]
    return np.asarray(idx, dtype=np.int64)


def sum_kbn(x):
    """KahanSummation.sum_kbn (Kahan-Babuska-Neumaier compensated sum)."""
    x = np.asarray(x, dtype=np.float64).ravel()
    if x.size == 0:
        return 0.0
    s = float(x[0])
    c = 0.0
    for xi in x[1:]:
        xi = float(xi)
        t = s + xi
        if abs(s) >= abs(xi):
            c += (s - t) + xi
        else:
            c += (xi - t) + s
        s = t
    return s + c


# ------------------------------------------------------------------------------------------------
# the ten full conditionals  (src/gibbs.jl:267-636), each a pure function of
# (state pieces, data, injected basic variates) -> dict of parameters and the draw
# ------------------------------------------------------------------------------------------------
def update_tau2(X, y, V, mu_prev, gamma_prev, u_prev, lam_prev, S_prev, g):
    """src/gibbs.jl:267-277.  g = one Gamma(shape, 1) variate; tau2 = scale / g
    (InverseGamma(a, s) == s / Gamma(a, 1))."""
    n = y.shape[0]
    r = y - mu_prev - X @ gamma_prev
    gw = gamma_prev - W_of(u_prev, lam_prev)
    scale = float(r @ r) / 2.0 + sum_kbn((gw ** 2 / 2.0) / S_prev)
    shape = n / 2.0 + V * (V + 1) / 4.0
    return dict(shape=shape, scale=scale, tau2=scale / g)


def chol_with_jitter(A):
    """The reference's jitter ladder (src/gibbs.jl:322-347): try Cholesky of the lower triangle,
    else add 1e-5 I and retry, else add a further 4e-5 I and retry, else fail."""
    A = np.array(A, dtype=np.float64)
    Al = np.tril(A) + np.tril(A, -1).T  # Hermitian(., :L)
    jit = 0
    for add in (0.0, 1e-5, 4e-5):
        if add:
            Al = Al + add * np.eye(Al.shape[0])
            jit += 1
        try:
            return np.linalg.cholesky(Al), Al, jit
        except np.linalg.LinAlgError:
            continue
    raise np.linalg.LinAlgError("Sigma^-1 not positive definite after jitter ladder")


def xi_from_w(w, upsilon):
    """update_xi (src/gibbs.jl:385-402): w<=0 -> 1, w>=1 -> 0 (no variate consumed),
    else Bernoulli(1-w) realised as [upsilon <= 1-w]; NaN w -> Bernoulli(0.5)."""
    if w <= 0:
        return 1.0
    if w >= 1:
        return 0.0
    if math.isnan(w):
        return 1.0 if upsilon <= 0.5 else 0.0
    return 1.0 if upsilon <= 1.0 - w else 0.0


def update_u_xi_node(k, V, tau2, u_prev, lam_prev, S_prev, gamma_prev, Delta_prev, M_prev,
                     upsilon, z, literal=True):
    """One node of update_u_xi! (src/gibbs.jl:293-371).  Jacobi: uses u_prev of every other node.
    literal=True evaluates the two (V-1)-dimensional normal densities like the reference
    (src/gibbs.jl:349-351); literal=False uses the R x R identity
        log(w_bot/w_top) = log(Delta/(1-Delta)) - 1/2 logdet M - 1/2 logdet Sigma^-1 + 1/2 b' Sigma b
    (b = U' H^-1 gamma_k / tau2), which is what the CUDA kernel computes."""
    others = [l for l in range(V) if l != k]
    U = (u_prev[:, others].T) * lam_prev[None, :]          # (V-1) x R
    idx = node_edge_indices(k, V)
    h = S_prev[idx]
    gk = gamma_prev[idx]
    Minv = np.linalg.inv(M_prev)
    Sig_inv = (U.T @ (U / h[:, None])) / tau2 + Minv
    L, Sig_inv_used, jit = chol_with_jitter(Sig_inv)
    b = (U.T @ (gk / h)) / tau2
    mu_t = np.linalg.solve(Sig_inv_used, b)
    if literal:
        cov0 = tau2 * h
        lt = math.log1p(-Delta_prev) if Delta_prev < 1 else -math.inf
        lb = math.log(Delta_prev) if Delta_prev > 0 else -math.inf
        ll_top = lt - 0.5 * (np.sum(np.log(2 * np.pi * cov0)) + np.sum(gk ** 2 / cov0))
        C1 = np.diag(cov0) + U @ M_prev @ U.T
        sign, ld = np.linalg.slogdet(C1)
        ll_bot = lb - 0.5 * (len(gk) * math.log(2 * np.pi) + ld + gk @ np.linalg.solve(C1, gk))
        log_odds = ll_bot - ll_top
    else:
        _, ldM = np.linalg.slogdet(M_prev)
        ldS = 2.0 * np.sum(np.log(np.diag(L)))
        log_odds = (math.log(Delta_prev) - math.log1p(-Delta_prev)) - 0.5 * ldM - 0.5 * ldS \
            + 0.5 * float(b @ mu_t)
    w = 1.0 / (1.0 + math.exp(log_odds)) if log_odds < 700 else 0.0
    xi = xi_from_w(w, upsilon)
    # inv(C.U) z with C.U = L'  (src/gibbs.jl:365)
    u_tmp = mu_t + np.linalg.solve(L.T, z)
    return dict(Sigma_inv=Sig_inv_used, chol=L, mu_t=mu_t, log_odds=log_odds, w=w, xi=xi,
                u=xi * u_tmp, jitter=jit)


def update_u_xi(V, tau2, u_prev, lam_prev, S_prev, gamma_prev, Delta_prev, M_prev, upsilon, Z,
                literal=True):
    """All nodes (src/gibbs.jl:295): upsilon (V,), Z (V, R)."""
    R = u_prev.shape[0]
    out = dict(u=np.zeros((R, V)), xi=np.zeros(V), nodes=[])
    for k in range(V):
        r = update_u_xi_node(k, V, tau2, u_prev, lam_prev, S_prev, gamma_prev, Delta_prev, M_prev,
                             upsilon[k], Z[k], literal)
        out["u"][:, k] = r["u"]
        out["xi"][k] = r["xi"]
        out["nodes"].append(r)
    return out


def update_gamma(X, y, tau2, u_new, lam_prev, S_prev, mu_prev, z1, z2):
    """src/gibbs.jl:420-438 (Bhattacharya-style draw through an n x n system, LU solve)."""
    n = X.shape[0]
    W = W_of(u_new, lam_prev)
    tau = math.sqrt(tau2)
    d1 = np.sqrt(tau2 * S_prev) * z1          # MvNormal(0, diag) == sqrt(diag) .* z
    d2 = z2
    Xt = X / tau
    a1 = (y - X @ W - mu_prev) / tau
    a3 = Xt @ d1 + d2
    G = (Xt * (tau2 * S_prev)[None, :]) @ Xt.T + np.eye(n)
    a4 = np.linalg.solve(G, a1 - a3)
    gamma = d1 + (tau2 * S_prev) * (Xt.T @ a4) + W
    return dict(W=W, G=G, a1=a1, a3=a3, a4=a4, gamma=gamma, delta1=d1)


def update_gamma_qform(X, y, tau2, u_new, lam_prev, S_prev, mu_prev, z):
    """The same conditional as update_gamma (src/gibbs.jl:420-438: gamma | rest ~ N(W + m, P^-1)) drawn through
    the q x q precision: P = (X'X + D^-1)/tau2 = L L', L w = X'(y - mu - XW)/tau2, L' beta = w + z, gamma = W + beta.
    The reference never forms P; this restates BASELINE.json's q-form so the CUDA q-form has a checker.
    Equivalence with the reference's draw is in law (mean m and covariance P^-1, see gamma_conditional_moments),
    not per injected normal."""
    W = W_of(u_new, lam_prev)
    P = (X.T @ X + np.diag(1.0 / S_prev)) / tau2
    L = np.linalg.cholesky(P)
    b = X.T @ ((y - mu_prev - X @ W) / tau2)
    w = np.linalg.solve(L, b)
    beta = np.linalg.solve(L.T, w + z)
    return dict(W=W, P=P, L=L, b=b, w=w, beta=beta, mean=W + np.linalg.solve(L.T, w), gamma=W + beta)


def gamma_conditional_moments(X, y, tau2, W, S_prev, mu_prev):
    """Mean and precision of gamma | rest: P = (X'X + D^-1)/tau2, m = W + P^-1 X'(y-mu-XW)/tau2.
    Used by the distributional tests."""
    P = (X.T @ X + np.diag(1.0 / S_prev)) / tau2
    m = W + np.linalg.solve(P, X.T @ (y - mu_prev - X @ W) / tau2)
    return m, P


# ---- GIG sampler (src/gig.jl:8-176) driven by an injected uniform stream -------------------------
class UniformStream:
    """Sequential source of U(0,1) values (list/array or callable)."""

    def __init__(self, values):
        self.values = values
        self.pos = 0

    def __call__(self):
        v = self.values[self.pos]
        self.pos += 1
        return float(v)


def gig_mode(lam, omega):
    """src/gig.jl:170-176."""
    if lam >= 1.0:
        return (math.sqrt((lam - 1.0) ** 2 + omega ** 2) + lam - 1.0) / omega
    return omega / (math.sqrt((1.0 - lam) ** 2 + omega ** 2) + (1.0 - lam))


def gig_rou_shift(lam, omega, alpha, unif):
    """src/gig.jl:44-78."""
    t = 0.5 * (lam - 1.0)
    s = 0.25 * omega
    xm = gig_mode(lam, omega)
    nc = t * math.log(xm) - s * (xm + 1.0 / xm)
    a = -(2.0 * (lam + 1.0) / omega + xm)
    b = 2.0 * (lam - 1.0) * xm / omega - 1.0
    c = xm
    p = b - a * a / 3.0
    q = 2.0 * a ** 3 / 27.0 - a * b / 3.0 + c
    fi = math.acos(-q / (2.0 * math.sqrt(-p ** 3 / 27.0)))
    fak = 2.0 * math.sqrt(-p / 3.0)
    y1 = fak * math.cos(fi / 3.0) - a / 3.0
    y2 = fak * math.cos(fi / 3.0 + 4.0 / 3.0 * math.pi) - a / 3.0
    uplus = (y1 - xm) * math.exp(t * math.log(y1) - s * (y1 + 1.0 / y1) - nc)
    uminus = (y2 - xm) * math.exp(t * math.log(y2) - s * (y2 + 1.0 / y2) - nc)
    while True:
        U = uminus + unif() * (uplus - uminus)
        Vv = unif()
        Xv = U / Vv + xm
        if Xv > 0.0 and math.log(Vv) <= t * math.log(Xv) - s * (Xv + 1.0 / Xv) - nc:
            return alpha * Xv


def gig_rou_noshift(lam, omega, alpha, unif):
    """src/gig.jl:80-100."""
    t = 0.5 * (lam - 1.0)
    s = 0.25 * omega
    xm = gig_mode(lam, omega)
    nc = t * math.log(xm) - s * (xm + 1.0 / xm)
    ym = ((lam + 1.0) + math.sqrt((lam + 1.0) ** 2 + omega ** 2)) / omega
    um = math.exp(0.5 * (lam + 1.0) * math.log(ym) - s * (ym + 1.0 / ym) - nc)
    while True:
        U = um * unif()
        Vv = unif()
        Xv = U / Vv
        if math.log(Vv) <= t * math.log(Xv) - s * (Xv + 1.0 / Xv) - nc:
            return alpha * Xv


def gig_concave(lam, omega, alpha, unif):
    """src/gig.jl:102-168 (only the lam > 0 arms are reachable with lam = 1/2)."""
    xm = gig_mode(lam, omega)
    x0 = omega / (1.0 - lam)
    k0 = math.exp((lam - 1.0) * math.log(xm) - 0.5 * omega * (xm + 1.0 / xm))
    A1 = k0 * x0
    if x0 >= 2.0 / omega:
        k1 = 0.0
        A2 = 0.0
        k2 = x0 ** (lam - 1.0)
        A3 = k2 * 2.0 * math.exp(-omega * x0 / 2.0) / omega
    else:
        k1 = math.exp(-omega)
        if lam == 0.0:
            A2 = k1 * math.log(2.0 / omega ** 2)
        else:
            A2 = k1 / lam * ((2.0 / omega) ** lam - x0 ** lam)
        k2 = (2.0 / omega) ** (lam - 1.0)
        A3 = k2 * 2.0 * math.exp(-1.0) / omega
    Atot = A1 + A2 + A3
    while True:
        Vv = Atot * unif()
        if Vv <= A1:
            Xv = x0 * Vv / A1
            hx = k0
        else:
            Vv -= A1
            if Vv <= A2:
                if lam == 0.0:
                    Xv = omega * math.exp(math.exp(omega) * Vv)
                    hx = k1 / Xv
                else:
                    Xv = (x0 ** lam + lam / k1 * Vv) ** (1.0 / lam)
                    hx = k1 * Xv ** (lam - 1.0)
            else:
                Vv -= A2
                a = x0 if x0 > 2.0 / omega else 2.0 / omega
                Xv = -2.0 / omega * math.log(math.exp(-omega / 2.0 * a) - omega / (2.0 * k2) * Vv)
                hx = k2 * math.exp(-omega / 2.0 * Xv)
        U = unif() * hx
        if math.log(U) <= (lam - 1.0) * math.log(Xv) - omega / 2.0 * (Xv + 1.0 / Xv):
            return alpha * Xv


def sample_gig(lam, chi, psi, unif, gamma_variate=None):
    """src/gig.jl:8-42 for lam >= 0.  Degenerate arms (chi or psi < 10 eps) take one injected
    Gamma(lam, 1) variate g and return g*psi/2 (resp. 1/(g*chi/2)) -- the reference's own
    (non-GIGrvg) scale convention, src/gig.jl:15-26."""
    if chi < EPS10:
        g = gamma_variate if gamma_variate is not None else unif()
        return g * (psi / 2.0), "degenerate_chi"
    if psi < EPS10:
        g = gamma_variate if gamma_variate is not None else unif()
        return 1.0 / (g * (chi / 2.0)), "degenerate_psi"
    alpha = math.sqrt(chi / psi)
    omega = math.sqrt(psi * chi)
    if lam > 2.0 or omega > 3.0:
        return gig_rou_shift(lam, omega, alpha, unif), "shift"
    if lam >= 1.0 - 2.25 * omega ** 2 or omega > 0.2:
        return gig_rou_noshift(lam, omega, alpha, unif), "noshift"
    if lam >= 0.0 and omega > 0.0:
        return gig_concave(lam, omega, alpha, unif), "concave"
    raise ValueError("sample_gig fell off the end (reference returns nothing)")


def update_D(gamma_new, u_new, lam_prev, tau2, theta_prev, uniforms):
    """src/gibbs.jl:454-458 + sample_rgig (116-118): S_j ~ GIG(1/2, psi=theta, chi=(gamma_j-W_j)^2/tau2).
    uniforms: (q, K) array, row j = the uniform stream of edge j."""
    W = W_of(u_new, lam_prev)
    chi = (gamma_new - W) ** 2 / tau2
    S = np.empty_like(chi)
    branch = []
    used = np.zeros(chi.shape[0], dtype=np.int64)
    for j in range(chi.shape[0]):
        st = UniformStream(uniforms[j])
        S[j], br = sample_gig(0.5, float(chi[j]), float(theta_prev), st)
        branch.append(br)
        used[j] = st.pos
    return dict(chi=chi, S=S, branch=branch, used=used)


def update_theta(S_new, zeta, iota, V, g):
    """src/gibbs.jl:476-479: Gamma(shape zeta + V(V+1)/2, scale 2/(2 iota + sum S)); g ~ Gamma(shape,1)."""
    shape = zeta + V * (V + 1) / 2.0
    scale = 2.0 / (2.0 * iota + sum_kbn(S_new))
    return dict(shape=shape, scale=scale, theta=g * scale)


def update_Delta(xi_new, a_delta, b_delta, ga, gb, upsilon=0.0):
    """src/gibbs.jl:496-499 + sample_Beta (130-140).  ga ~ Gamma(a,1), gb ~ Gamma(b,1)."""
    a = a_delta + sum_kbn(xi_new)
    b = b_delta + sum_kbn(1.0 - xi_new)
    if a > 0.0 and b > 0.0:
        d = ga / (ga + gb)
    elif a > 0.0:
        d = 1.0
    elif b > 0.0:
        d = 0.0
    else:
        d = 0.0 if upsilon < 0.5 else 1.0
    return dict(a=a, b=b, Delta=d)


def update_M(u_new, xi_new, nu, c, zl):
    """src/gibbs.jl:516-547: InverseWishart(nu + #{xi != 0}, Psi = I + sum_k u_k u_k').
    Realisation (distributionally identical to Distributions.jl's; variate algorithm unpinned):
    Bartlett factor A (lower; A_ii = sqrt(c_i), c_i ~ chi2(df - i), i = 0..R-1; A_ij = zl[i(i-1)/2+j],
    j < i), Psi = Lp Lp', T = Lp A^-T, M = T T'."""
    R = u_new.shape[0]
    Psi = np.eye(R) + u_new @ u_new.T
    df = nu + int(np.sum(np.abs(xi_new) > 0.1))
    Lp = np.linalg.cholesky(Psi)
    A = np.zeros((R, R))
    t = 0
    for i in range(R):
        A[i, i] = math.sqrt(c[i])
        for j in range(i):
            A[i, j] = zl[t]
            t += 1
    T = np.linalg.solve(A, Lp.T).T     # Lp A^-T
    return dict(df=df, Psi=Psi, chol_Psi=Lp, M=T @ T.T)


def update_mu(Xgamma_new, y, tau2, z):
    """src/gibbs.jl:565-570."""
    n = y.shape[0]
    m = float(np.mean(y - Xgamma_new))
    sd = math.sqrt(tau2 / n)
    return dict(mean=m, sd=sd, mu=m + sd * z)


LAMBDA_VALUES = (0.0, 1.0, -1.0)


def categorical(weights, upsilon):
    """StatsBase.sample(rng, values, weights) [memory]: t = upsilon * sum(w); first i with cumsum >= t."""
    t = upsilon * float(np.sum(weights))
    cw = float(weights[0])
    i = 0
    while cw < t and i < len(weights) - 1:
        i += 1
        cw += float(weights[i])
    return i


def update_lambda(gamma_new, u_new, S_new, tau2, lam_prev, pi_prev, upsilon):
    """src/gibbs.jl:586-613.  Every r uses lam_prev for the other components (Lambda is built once)."""
    R = lam_prev.shape[0]
    sd2 = tau2 * S_new
    lam_new = np.empty(R)
    ll = np.empty((R, 3))
    wts = np.empty((R, 3))
    for r in range(R):
        for c, v in enumerate(LAMBDA_VALUES):
            l2 = lam_prev.copy()
            l2[r] = v
            Wv = W_of(u_new, l2)
            ll[r, c] = sum_kbn(-0.5 * np.log(2 * np.pi * sd2) - 0.5 * (gamma_new - Wv) ** 2 / sd2)
        a1 = np.exp(ll[r] - ll[r].max())
        wts[r] = pi_prev[r] * a1
        lam_new[r] = LAMBDA_VALUES[categorical(wts[r], upsilon[r])]
    return dict(loglik=ll, weights=wts, lam=lam_new)


def update_pi(lam_new, eta, g):
    """src/gibbs.jl:630-636 + sample_pi_dirichlet! (159-169).  g: (R,3) unit-scale gamma variates
    with shapes alpha[r]; pi[r] = g[r]/sum(g[r])."""
    R = lam_new.shape[0]
    alpha = np.empty((R, 3))
    for r in range(R):
        alpha[r] = [(r + 1) ** eta, 1.0, 1.0]
        if lam_new[r] == 1:
            alpha[r, 1] += 1
        elif lam_new[r] == 0:
            alpha[r, 0] += 1
        else:
            alpha[r, 2] += 1
    return dict(alpha=alpha, pi=g / g.sum(axis=1, keepdims=True))


# ------------------------------------------------------------------------------------------------
# R-hat and Summary  (src/convergence.jl:4-65, src/gibbs.jl:1214-1250) -- pinned exactly by goldens
# ------------------------------------------------------------------------------------------------
def rhat(chains):
    """chains: (niter, nparams, nchains).  Split every chain in two halves of niter//2 (first half
    = draws 0..h-1, second half = the LAST h draws, so for odd niter the middle draw is dropped,
    MCMCDiagnosticTools 0.1.x copyto_split! [memory])."""
    niter_full, nparams, nch = chains.shape
    h = niter_full // 2
    if h - 1 <= 0:
        return np.full(nparams, np.nan)
    first = chains[:h]
    second = chains[niter_full - h:]
    samples = np.concatenate([first, second], axis=2)      # (h, nparams, 2*nch)
    cm = samples.mean(axis=0)                               # (nparams, 2nch)
    cv = ((samples - cm[None]) ** 2).sum(axis=0) / (h - 1)
    W = cv.mean(axis=1)
    varp = (h - 1) / h * W + cm.var(axis=1, ddof=1)
    out = np.empty(nparams)
    for i in range(nparams):
        if varp[i] == 0 and W[i] == 0:
            out[i] = 1.0
        elif W[i] == 0:
            out[i] = np.inf
        else:
            out[i] = math.sqrt(varp[i] / W[i])
    return out


def rhat_from_moments(mean, m2, h):
    """Same estimator from per-(split-chain) means and sums of squared deviations.
    mean, m2: (nsplit, nparams); h = draws per split chain."""
    W = (m2 / (h - 1)).mean(axis=0)
    varp = (h - 1) / h * W + mean.var(axis=0, ddof=1)
    out = np.where((varp == 0) & (W == 0), 1.0, np.where(W == 0, np.inf, np.sqrt(varp / np.where(W == 0, 1, W))))
    return out


def ess_geyer(x, max_lag=None):
    """Multi-chain effective sample size of one scalar parameter; x: (draws, chains).
    NOT in the reference (BayesianNetworkRegression.jl computes no ESS; BASELINE.json asks for "gamma ESS/sec").
    Estimator (Geyer 1992 initial monotone sequence, multi-chain form of Vehtari et al. 2021 / Stan, without
    rank-normalisation, as MCMCDiagnosticTools' ess with the default estimator):
      acov_c(t) biased per-chain autocovariance, W = mean_c acov_c(0) n/(n-1), var+ = W (n-1)/n + var(chain means),
      rho_t = 1 - (W - mean_c acov_c(t)) / var+,  P_k = rho_2k + rho_2k+1 truncated at the first negative pair and
      made non-increasing,  tau = -1 + 2 sum P_k,  ESS = n m / max(tau, 1/log10(n m)).
    max_lag (odd) bounds the lags that may be used, mirroring the device kernel's lag budget."""
    x = np.asarray(x, dtype=np.float64)
    n, m = x.shape
    xc = x - x.mean(axis=0, keepdims=True)
    L = n - 1 if max_lag is None else min(int(max_lag), n - 1)
    acov = np.stack([(xc[: n - t] * xc[t:]).sum(axis=0) / n for t in range(L + 1)])      # (L+1, m), direct lags
    W = acov[0].mean() * n / (n - 1)
    B_over_n = x.mean(axis=0).var(ddof=1) if m > 1 else 0.0
    var_plus = W * (n - 1) / n + B_over_n
    if not var_plus > 0:
        return float("nan")
    rho = 1.0 - (W - acov.mean(axis=1)) / var_plus
    tau, prev, t = -1.0, np.inf, 0
    while t + 1 <= L and t + 1 < n:
        pair = rho[t] + rho[t + 1]
        if pair < 0:
            break
        pair = min(pair, prev)
        tau += 2.0 * pair
        prev = pair
        t += 2
    nm = n * m
    return float(nm / max(tau, 1.0 / math.log10(max(nm, 10))))


def ess_stats(x, max_lag):
    """Per-rank sufficient statistics of ess_geyer for x (draws, chains): sum over the chains of the biased
    per-chain-centred autocovariances, lags 0..max_lag, and the chain means -- the two buffers of bnr_ess_device."""
    n = x.shape[0]
    xc = x - x.mean(axis=0, keepdims=True)
    acov = np.stack([(xc[: n - t] * xc[t:]).sum(axis=0) / n for t in range(max_lag + 1)])
    return acov.sum(axis=1), x.mean(axis=0)


def ess_from_stats(acov_parts, chain_means, n, max_lag):
    """ess_geyer from gathered statistics (bnr_ess_from_stats): acov_parts (ranks, max_lag+1), chain_means (chains,)."""
    m = len(chain_means)
    acov = np.asarray(acov_parts).sum(axis=0) / m
    W = acov[0] * n / (n - 1)
    B_over_n = np.var(chain_means, ddof=1) if m > 1 else 0.0
    var_plus = W * (n - 1) / n + B_over_n
    if not var_plus > 0:
        return float("nan")
    rho = 1.0 - (W - acov) / var_plus
    tau, prev, t = -1.0, np.inf, 0
    while t + 1 <= max_lag and t + 1 < n:
        pair = rho[t] + rho[t + 1]
        if pair < 0:
            break
        pair = min(pair, prev)
        tau += 2.0 * pair
        prev = pair
        t += 2
    nm = n * m
    return float(nm / max(tau, 1.0 / math.log10(max(nm, 10))))


def julia_round(x):
    """Julia round(): ties to even."""
    return int(np.rint(x))


def summary(gamma_trace, xi_trace, interval=95, digits=3):
    """src/gibbs.jl:1214-1250.  gamma_trace (nsamp, q), xi_trace (nsamp, V): post-burn rows of chain 1."""
    nsamp, q = gamma_trace.shape
    lower = (100 - interval) / 200.0
    lw = julia_round(nsamp * lower)
    hi = julia_round(nsamp * (1.0 - lower))
    if lw < 1:
        raise IndexError("nsamp too small for the requested interval (reference raises BoundsError)")
    gs = np.sort(gamma_trace, axis=0)
    V = int((-1 + math.sqrt(1 + 8 * q)) / 2)
    node1 = np.concatenate([np.full(V - k, k + 1) for k in range(V)])
    node2 = np.concatenate([np.arange(k + 1, V + 1) for k in range(V)])
    return dict(node1=node1, node2=node2,
                estimate=np.round(gamma_trace.mean(axis=0), digits),
                lower_bound=np.round(gs[lw - 1], digits),
                upper_bound=np.round(gs[hi - 1], digits),
                probability=np.round(xi_trace.mean(axis=0), digits),
                ci_level=interval)


# ------------------------------------------------------------------------------------------------
# one full sweep from injected basic variates  (src/gibbs.jl:663-677) and the prior init (191-224)
# ------------------------------------------------------------------------------------------------
def draw_layout(n, V, R, K_gig):
    """Offsets of every draw site inside one chain's injected-variate vector (shared with the CUDA
    library's bnr_set_injection; see include/bnr.h)."""
    q = V * (V + 1) // 2
    sizes = [("tau2", 1), ("uxi", V * (R + 1)), ("gamma_z1", q), ("gamma_z2", n), ("S", q * K_gig),
             ("theta", 1), ("Delta", 3), ("M", R + R * (R - 1) // 2), ("mu", 1), ("lambda", R),
             ("pi", 3 * R)]
    off, o = {}, 0
    for name, s in sizes:
        off[name] = (o, s)
        o += s
    off["_total"] = o
    return off


def gibbs_sweep(state, X, y, V, R, hyper, inj, K_gig, literal=True, gamma_form="n"):
    """One gibbs_sample! with every basic variate read from `inj` (layout: draw_layout).
    gamma_form "n" = the reference's n x n draw (z1, z2), "q" = the q x q precision draw (z1 only).
    Returns (new_state, aux)."""
    n = X.shape[0]
    lay = draw_layout(n, V, R, K_gig)

    def seg(name):
        o, s = lay[name]
        return inj[o:o + s]

    aux = {}
    new = dict(state)
    t = update_tau2(X, y, V, state["mu"], state["gamma"], state["u"], state["lam"], state["S"], seg("tau2")[0])
    new["tau2"] = t["tau2"]
    aux["tau2"] = t
    uz = seg("uxi").reshape(V, R + 1)
    ux = update_u_xi(V, new["tau2"], state["u"], state["lam"], state["S"], state["gamma"], state["Delta"],
                     state["M"], uz[:, 0], uz[:, 1:], literal)
    new["u"], new["xi"] = ux["u"], ux["xi"]
    aux["uxi"] = ux
    if gamma_form == "q":
        g = update_gamma_qform(X, y, new["tau2"], new["u"], state["lam"], state["S"], state["mu"], seg("gamma_z1"))
    else:
        g = update_gamma(X, y, new["tau2"], new["u"], state["lam"], state["S"], state["mu"],
                         seg("gamma_z1"), seg("gamma_z2"))
    new["gamma"] = g["gamma"]
    aux["gamma"] = g
    q = V * (V + 1) // 2
    d = update_D(new["gamma"], new["u"], state["lam"], new["tau2"], state["theta"], seg("S").reshape(q, K_gig))
    new["S"] = d["S"]
    aux["S"] = d
    th = update_theta(new["S"], hyper["zeta"], hyper["iota"], V, seg("theta")[0])
    new["theta"] = th["theta"]
    aux["theta"] = th
    dd = seg("Delta")
    de = update_Delta(new["xi"], hyper["a_delta"], hyper["b_delta"], dd[0], dd[1], dd[2])
    new["Delta"] = de["Delta"]
    aux["Delta"] = de
    mm = seg("M")
    m = update_M(new["u"], new["xi"], hyper["nu"], mm[:R], mm[R:])
    new["M"] = m["M"]
    aux["M"] = m
    mu = update_mu(X @ new["gamma"], y, new["tau2"], seg("mu")[0])
    new["mu"] = mu["mu"]
    aux["mu"] = mu
    la = update_lambda(new["gamma"], new["u"], new["S"], new["tau2"], state["lam"], state["pi"], seg("lambda"))
    new["lam"] = la["lam"]
    aux["lambda"] = la
    p = update_pi(new["lam"], hyper["eta"], seg("pi").reshape(R, 3))
    new["pi"] = p["pi"]
    aux["pi"] = p
    return new, aux


DEFAULT_HYPER = dict(eta=1.01, zeta=1.0, iota=1.0, a_delta=1.0, b_delta=1.0, nu=10)


def init_layout(V, R):
    """Basic variates consumed by initialize_variables! in the reference's order (src/gibbs.jl:199-223)."""
    q = V * (V + 1) // 2
    sizes = [("S", q), ("pi", 3 * R), ("lambda", R), ("xi", V), ("M", R + R * (R - 1) // 2),
             ("u", V * R), ("gamma", q)]
    off, o = {}, 0
    for name, s in sizes:
        off[name] = (o, s)
        o += s
    off["_total"] = o
    return off


def initialize_state(V, R, hyper, inj):
    """src/gibbs.jl:191-224 from injected basic variates (layout: init_layout):
    S: unit exponentials e -> S = e*theta/2; pi: Gamma(alpha,1) variates; lambda, xi: uniforms;
    M: Bartlett variates (as update_M with Psi = I, df = nu); u, gamma: standard normals."""
    lay = init_layout(V, R)
    q = V * (V + 1) // 2

    def seg(name):
        o, s = lay[name]
        return inj[o:o + s]

    eta = hyper["eta"] if hyper["eta"] > 1 else 1.01
    st = {}
    st["theta"] = 0.5
    st["S"] = seg("S") * (st["theta"] / 2.0)
    g = seg("pi").reshape(R, 3)
    st["pi"] = g / g.sum(axis=1, keepdims=True)
    ul = seg("lambda")
    st["lam"] = np.array([LAMBDA_VALUES[categorical(st["pi"][r], ul[r])] for r in range(R)])
    st["Delta"] = 0.5
    st["xi"] = (seg("xi") <= st["Delta"]).astype(np.float64)
    mm = seg("M")
    A = np.zeros((R, R))
    t = 0
    for i in range(R):
        A[i, i] = math.sqrt(mm[i])
        for j in range(i):
            A[i, j] = mm[R + t]
            t += 1
    T = np.linalg.inv(A).T
    st["M"] = T @ T.T
    st["u"] = seg("u").reshape(V, R).T.copy()
    st["mu"] = 1.0
    st["tau2"] = 1.0
    W = W_of(st["u"], st["lam"])
    st["gamma"] = W + np.sqrt(st["tau2"] * st["S"]) * seg("gamma")
    return st
