"""CPU baseline: a reference-faithful TIMING port of gibbs_sample!  --  bench infrastructure only.

Julia is not installed here or on the GPU box, so the reference's own `Fit!` cannot be timed.  This port keeps
the reference's *formulation and BLAS/LAPACK calls* (src/gibbs.jl:267-636): the full (non-SYRK) GEMM
`X tau2D X'` followed by an LU solve for gamma, a dense (V-1)-dimensional covariance Cholesky per node for the
xi odds, three W rebuilds per latent dimension for lambda -- and replaces only what Julia would compile to
tight scalar loops (GIG rejection loops, element-wise maps, sum_kbn) by vectorised NumPy, so that Python
interpreter overhead is not charged to the reference.  One chain per process, OPENBLAS/MKL/OpenMP threads pinned to
1 *before NumPy is loaded* (the timing runs in a child interpreter, every worker asserts a 1-thread pool),
min(num_chains, cores) processes, exactly like one chain per Distributed.jl worker (src/gibbs.jl:946-948).
It is labelled kind="port" in bench.py.  Its draws are validated against the conditionals of bnr_oracle.py in
tests/test_cpu_baseline.py.
"""
import math
import os
import time

import numpy as np


def lower_triangle_idx(V):
    k, l = [], []
    for kk in range(V):
        for ll in range(kk, V):
            k.append(kk)
            l.append(ll)
    return np.asarray(l), np.asarray(k)


def gig_half_vec(chi, psi, rng):
    """Vectorised src/gig.jl for lambda = 1/2 (same three algorithms, same constants)."""
    lam = 0.5
    out = np.empty_like(chi)
    alpha = np.sqrt(chi / psi)
    omega = np.sqrt(psi * chi)
    shift = omega > 3.0
    nosh = ~shift & ((lam >= 1.0 - 2.25 * omega ** 2) | (omega > 0.2))
    conc = ~shift & ~nosh

    def mode(om):
        return om / (np.sqrt((1.0 - lam) ** 2 + om ** 2) + (1.0 - lam))

    def rejection(idx, propose, accept):
        todo = np.arange(idx.size)
        res = np.empty(idx.size)
        while todo.size:
            X, aux = propose(todo)
            ok = accept(todo, X, aux)
            res[todo[ok]] = X[ok]
            todo = todo[~ok]
        return res

    if nosh.any():
        om = omega[nosh]
        t, s = 0.5 * (lam - 1.0), 0.25 * om
        xm = mode(om)
        nc = t * np.log(xm) - s * (xm + 1.0 / xm)
        ym = ((lam + 1.0) + np.sqrt((lam + 1.0) ** 2 + om ** 2)) / om
        um = np.exp(0.5 * (lam + 1.0) * np.log(ym) - s * (ym + 1.0 / ym) - nc)

        def prop(td):
            U = um[td] * rng.random(td.size)
            Vv = rng.random(td.size)
            return U / Vv, Vv

        def acc(td, X, Vv):
            return np.log(Vv) <= t * np.log(X) - s[td] * (X + 1.0 / X) - nc[td]

        out[nosh] = alpha[nosh] * rejection(om, prop, acc)
    if shift.any():
        om = omega[shift]
        t, s = 0.5 * (lam - 1.0), 0.25 * om
        xm = mode(om)
        nc = t * np.log(xm) - s * (xm + 1.0 / xm)
        a = -(2.0 * (lam + 1.0) / om + xm)
        b = 2.0 * (lam - 1.0) * xm / om - 1.0
        c = xm
        p = b - a * a / 3.0
        q = 2.0 * a ** 3 / 27.0 - a * b / 3.0 + c
        fi = np.arccos(-q / (2.0 * np.sqrt(-p ** 3 / 27.0)))
        fak = 2.0 * np.sqrt(-p / 3.0)
        y1 = fak * np.cos(fi / 3.0) - a / 3.0
        y2 = fak * np.cos(fi / 3.0 + 4.0 / 3.0 * math.pi) - a / 3.0
        up = (y1 - xm) * np.exp(t * np.log(y1) - s * (y1 + 1.0 / y1) - nc)
        um = (y2 - xm) * np.exp(t * np.log(y2) - s * (y2 + 1.0 / y2) - nc)

        def prop(td):
            U = um[td] + rng.random(td.size) * (up[td] - um[td])
            Vv = rng.random(td.size)
            return U / Vv + xm[td], Vv

        def acc(td, X, Vv):
            with np.errstate(invalid="ignore", divide="ignore"):
                return (X > 0) & (np.log(Vv) <= t * np.log(np.abs(X)) - s[td] * (X + 1.0 / X) - nc[td])

        out[shift] = alpha[shift] * rejection(om, prop, acc)
    if conc.any():
        om = omega[conc]
        xm = mode(om)
        x0 = om / (1.0 - lam)
        k0 = np.exp((lam - 1.0) * np.log(xm) - 0.5 * om * (xm + 1.0 / xm))
        A1 = k0 * x0
        big = x0 >= 2.0 / om
        k1 = np.where(big, 0.0, np.exp(-om))
        A2 = np.where(big, 0.0, k1 / lam * ((2.0 / om) ** lam - x0 ** lam))
        k2 = np.where(big, x0 ** (lam - 1.0), (2.0 / om) ** (lam - 1.0))
        A3 = np.where(big, k2 * 2.0 * np.exp(-om * x0 / 2.0) / om, k2 * 2.0 * math.exp(-1.0) / om)
        At = A1 + A2 + A3

        def prop(td):
            Vv = At[td] * rng.random(td.size)
            o, x0t, A1t, A2t, k0t, k1t, k2t = om[td], x0[td], A1[td], A2[td], k0[td], k1[td], k2[td]
            r1 = Vv <= A1t
            V2 = Vv - A1t
            r2 = ~r1 & (V2 <= A2t)
            V3 = V2 - A2t
            with np.errstate(all="ignore"):
                X1 = x0t * Vv / A1t
                X2 = (x0t ** lam + lam / np.where(k1t > 0, k1t, 1.0) * V2) ** (1.0 / lam)
                a3 = np.maximum(x0t, 2.0 / o)
                X3 = -2.0 / o * np.log(np.exp(-o / 2.0 * a3) - o / (2.0 * k2t) * V3)
                X = np.where(r1, X1, np.where(r2, X2, X3))
                hx = np.where(r1, k0t, np.where(r2, k1t * X ** (lam - 1.0), k2t * np.exp(-o / 2.0 * X)))
            return X, hx

        def acc(td, X, hx):
            U = rng.random(td.size) * hx
            with np.errstate(all="ignore"):
                return np.log(U) <= (lam - 1.0) * np.log(X) - om[td] / 2.0 * (X + 1.0 / X)

        out[conc] = alpha[conc] * rejection(om, prop, acc)
    return out


class ReferencePort:
    """State + one-sweep method with the reference's dense formulation."""

    def __init__(self, X, y, R, seed, hyper=None):
        self.X = np.ascontiguousarray(X, dtype=np.float64)
        self.y = np.asarray(y, dtype=np.float64)
        self.n, self.q = self.X.shape
        self.V = int((-1 + math.sqrt(1 + 8 * self.q)) / 2)
        self.R = R
        self.h = dict(eta=1.01, zeta=1.0, iota=1.0, a_delta=1.0, b_delta=1.0, nu=10)
        self.h.update(hyper or {})
        self.rng = np.random.default_rng(seed)
        self.il, self.ik = lower_triangle_idx(self.V)
        self._init()

    def W(self, u, lam):
        full = u.T @ (lam[:, None] * u)          # transpose(u) * Diagonal(lam) * u
        return full[self.il, self.ik]

    def _init(self):
        rng, V, R, q, h = self.rng, self.V, self.R, self.q, self.h
        st = {}
        st["theta"] = 0.5
        st["S"] = rng.exponential(st["theta"] / 2.0, size=q)
        st["pi"] = np.stack([rng.dirichlet([(r + 1) ** h["eta"], 1, 1]) for r in range(R)])
        st["lam"] = np.array([rng.choice([0.0, 1.0, -1.0], p=st["pi"][r]) for r in range(R)])
        st["Delta"] = 0.5
        st["xi"] = (rng.random(V) <= 0.5).astype(float)
        A = np.zeros((R, R))
        for i in range(R):
            A[i, i] = math.sqrt(rng.chisquare(h["nu"] - i))
            A[i, :i] = rng.standard_normal(i)
        T = np.linalg.inv(A).T
        st["M"] = T @ T.T
        st["u"] = rng.standard_normal((R, V))
        st["mu"], st["tau2"] = 1.0, 1.0
        st["gamma"] = self.W(st["u"], st["lam"]) + np.sqrt(st["S"]) * rng.standard_normal(q)
        self.st = st

    def sweep(self):
        X, y, rng, V, R, q, n, h = self.X, self.y, self.rng, self.V, self.R, self.q, self.n, self.h
        s = self.st
        new = {}
        # tau2 (267-277)
        r = y - s["mu"] - X @ s["gamma"]
        gw = s["gamma"] - self.W(s["u"], s["lam"])
        scale = (r @ r) / 2 + np.sum((gw ** 2 / 2) / s["S"])
        tau2 = scale / rng.gamma(n / 2 + V * (V + 1) / 4)
        new["tau2"] = tau2
        # u, xi (293-371): dense per-node formulation
        Smat = np.zeros((V, V)); Smat[self.il, self.ik] = s["S"]; Smat = Smat + np.tril(Smat, -1).T
        Gmat = np.zeros((V, V)); Gmat[self.il, self.ik] = s["gamma"]; Gmat = Gmat + np.tril(Gmat, -1).T
        Minv = np.linalg.inv(s["M"])
        u_new = np.zeros((R, V)); xi_new = np.zeros(V)
        lt, lb = math.log1p(-s["Delta"]), math.log(s["Delta"])
        for k in range(V):
            oth = np.r_[0:k, k + 1:V]
            U = s["u"][:, oth].T * s["lam"][None, :]
            hk = Smat[k, oth]; gk = Gmat[k, oth]
            Sig_inv = (U.T @ (U / hk[:, None])) / tau2 + Minv
            L = np.linalg.cholesky(Sig_inv)
            cov0 = tau2 * hk
            ll_top = lt - 0.5 * (np.sum(np.log(2 * np.pi * cov0)) + np.sum(gk ** 2 / cov0))
            C1 = np.diag(cov0) + U @ s["M"] @ U.T          # dense (V-1) x (V-1), as the reference builds it
            Lc = np.linalg.cholesky(C1)
            zc = np.linalg.solve(Lc, gk)
            ll_bot = lb - 0.5 * (len(gk) * math.log(2 * np.pi) + 2 * np.sum(np.log(np.diag(Lc))) + zc @ zc)
            d = ll_bot - ll_top
            w = 1.0 / (1.0 + math.exp(d)) if d < 700 else 0.0
            xi = 1.0 if rng.random() <= 1 - w else 0.0
            mu_t = np.linalg.solve(Sig_inv, (U.T @ (gk / hk)) / tau2)
            u_new[:, k] = xi * (mu_t + np.linalg.solve(L.T, rng.standard_normal(R)))
            xi_new[k] = xi
        new["u"], new["xi"] = u_new, xi_new
        # gamma (420-438): full GEMM + LU solve
        Wn = self.W(u_new, s["lam"])
        tau = math.sqrt(tau2)
        d1 = np.sqrt(tau2 * s["S"]) * rng.standard_normal(q)
        d2 = rng.standard_normal(n)
        Xt = X / tau
        a1 = (y - X @ Wn - s["mu"]) / tau
        a3 = Xt @ d1 + d2
        G = (Xt * (tau2 * s["S"])[None, :]) @ Xt.T + np.eye(n)
        a4 = np.linalg.solve(G, a1 - a3)
        gamma = d1 + (tau2 * s["S"]) * (Xt.T @ a4) + Wn
        new["gamma"] = gamma
        # D (454-458)
        chi = (gamma - Wn) ** 2 / tau2
        new["S"] = gig_half_vec(np.maximum(chi, 1e-300), s["theta"], rng)
        # theta, Delta, M, mu
        new["theta"] = rng.gamma(h["zeta"] + q) * 2 / (2 * h["iota"] + new["S"].sum())
        new["Delta"] = rng.beta(h["a_delta"] + xi_new.sum(), h["b_delta"] + V - xi_new.sum())
        Psi = np.eye(R) + u_new @ u_new.T
        df = h["nu"] + int(xi_new.sum())
        Lp = np.linalg.cholesky(Psi)
        A = np.zeros((R, R))
        for i in range(R):
            A[i, i] = math.sqrt(rng.chisquare(df - i))
            A[i, :i] = rng.standard_normal(i)
        T = np.linalg.solve(A, Lp.T).T
        new["M"] = T @ T.T
        new["mu"] = np.mean(y - X @ gamma) + math.sqrt(tau2 / n) * rng.standard_normal()
        # lambda (586-613): three W rebuilds per r
        sd2 = tau2 * new["S"]
        lam_new = np.empty(R)
        const = -0.5 * np.log(2 * np.pi * sd2)
        for r_ in range(R):
            ll = np.empty(3)
            for c, v in enumerate((0.0, 1.0, -1.0)):
                l2 = s["lam"].copy(); l2[r_] = v
                ll[c] = np.sum(const - 0.5 * (gamma - self.W(u_new, l2)) ** 2 / sd2)
            wts = s["pi"][r_] * np.exp(ll - ll.max())
            lam_new[r_] = (0.0, 1.0, -1.0)[rng.choice(3, p=wts / wts.sum())]
        new["lam"] = lam_new
        al = np.array([[(r_ + 1) ** h["eta"], 1.0, 1.0] for r_ in range(R)])
        for r_ in range(R):
            al[r_, {0.0: 0, 1.0: 1, -1.0: 2}[lam_new[r_]]] += 1
        g = rng.gamma(al)
        new["pi"] = g / g.sum(axis=1, keepdims=True)
        self.st = new
        return new


PIN_VARS = ("OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "OMP_NUM_THREADS", "NUMEXPR_NUM_THREADS",
            "VECLIB_MAXIMUM_THREADS", "BLIS_NUM_THREADS")


def blas_threads():
    """Largest thread-pool size of any BLAS/OpenMP runtime loaded into this process (threadpoolctl)."""
    from threadpoolctl import threadpool_info
    return max([int(p.get("num_threads", 1)) for p in threadpool_info()] or [1])


def _worker(args):
    """One chain = one process (src/gibbs.jl:946-948).  The pool is forked from an interpreter whose BLAS was loaded
    with *_NUM_THREADS=1 (see time_port); this is asserted here, inside the worker, before anything is timed."""
    X, y, R, seed, warm, sweeps, barrier = args
    from threadpoolctl import threadpool_limits
    threadpool_limits(1)                       # belt and braces: also caps a pool that was sized before the fork
    nthr = blas_threads()
    if nthr != 1:
        raise RuntimeError("CPU baseline worker has a %d-thread BLAS pool; it must be 1" % nthr)
    port = ReferencePort(X, y, R, seed)
    for _ in range(warm):
        port.sweep()
    barrier.wait()                             # every chain has warmed up: the timed region starts together
    t0 = time.perf_counter()
    for _ in range(sweeps):
        port.sweep()
    t1 = time.perf_counter()
    barrier.wait()
    return t0, t1, nthr


def host_cores():
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    return cores


def _time_port_here(X, y, R, nproc, sweeps, warm):
    """Runs in an interpreter that was STARTED with the BLAS pools pinned to one thread."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    mgr_barrier = ctx.Barrier(nproc)
    procs, q = [], ctx.Queue()

    def run(c):
        try:
            q.put((c, _worker((X, y, R, 1000 + c, warm, sweeps, mgr_barrier))))
        except BaseException as exc:          # a dead worker must not leave the others at the barrier
            mgr_barrier.abort()
            q.put((c, exc))

    for c in range(nproc):
        p = ctx.Process(target=run, args=(c,))
        p.start()
        procs.append(p)
    res = [q.get() for _ in range(nproc)]
    for p in procs:
        p.join()
    for _, r in res:
        if isinstance(r, BaseException):
            raise r
    t0 = min(r[0] for _, r in res)
    t1 = max(r[1] for _, r in res)
    return dict(timed_s=t1 - t0, threads_per_proc=max(r[2] for _, r in res), nproc=nproc, sweeps=sweeps, warm=warm)


def time_port(X, y, R, chains, sweeps, warm=1, nproc=None):
    """Time `sweeps` Gibbs sweeps of min(chains, cores) independent chains, one single-threaded chain per process,
    all concurrently (one chain per Distributed.jl worker, src/gibbs.jl:946-948).

    The BLAS thread pools are sized when NumPy is first imported, so setting *_NUM_THREADS afterwards (or in a forked
    child) does nothing: the measurement therefore runs in a CHILD INTERPRETER started with the variables already in
    its environment, every worker asserts a 1-thread pool, and the timed region is bracketed by barriers (start-up,
    data transfer and warm-up are outside it).  Returns a dict: value (chain-iterations/s summed over the chains),
    nproc, threads_per_proc, sweeps, warm, timed_s, wall_s."""
    import json
    import subprocess
    import sys
    import tempfile
    nproc = min(chains, host_cores()) if nproc is None else nproc
    env = dict(os.environ)
    for var in PIN_VARS:
        env[var] = "1"
    t0 = time.perf_counter()
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "xy.npz")
        np.savez(path, X=np.ascontiguousarray(X), y=np.asarray(y))
        cmd = [sys.executable, os.path.abspath(__file__), path, str(int(R)), str(int(nproc)), str(int(sweeps)),
               str(int(warm))]
        out = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("CPU baseline child failed:\n" + out.stderr[-2000:])
    r = json.loads(out.stdout.strip().splitlines()[-1])
    r["wall_s"] = time.perf_counter() - t0
    r["value"] = r["nproc"] * r["sweeps"] / r["timed_s"]
    return r


if __name__ == "__main__":
    import json
    import sys
    _path, _R, _nproc, _sweeps, _warm = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    if blas_threads() != 1:
        raise SystemExit("BLAS pool of the timing interpreter is not single-threaded: %d" % blas_threads())
    _d = np.load(_path)
    print(json.dumps(_time_port_here(_d["X"], _d["y"], _R, _nproc, _sweeps, _warm)))
