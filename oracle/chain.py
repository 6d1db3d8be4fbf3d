"""Self-driven oracle chain -- TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/bnr_oracle.py header).

Runs the restated Gibbs sweep on the CPU with NumPy's generator supplying the basic variates that the
injected-draw conditionals expect.  Used (a) as the level-2 posterior reference in tests and (b) as the
`cpu_baseline` / `--impl reference` timing leg of bench.py (kind "port": Julia is not installed, so the
reference's own Fit! cannot be timed).
"""
import math

import numpy as np

from . import bnr_oracle as O


def sweep_variates(rng, n, V, R, K, state, hyper):
    """Basic variates of one sweep in draw_layout order.  Gamma shapes follow the conditionals."""
    lay = O.draw_layout(n, V, R, K)
    q = V * (V + 1) // 2
    inj = np.empty(lay["_total"])

    def put(name, vals):
        o, s = lay[name]
        inj[o:o + s] = np.asarray(vals).ravel()

    put("tau2", rng.gamma(n / 2.0 + V * (V + 1) / 4.0))
    uz = np.empty((V, R + 1))
    uz[:, 0] = rng.random(V)
    uz[:, 1:] = rng.standard_normal((V, R))
    put("uxi", uz)
    put("gamma_z1", rng.standard_normal(q))
    put("gamma_z2", rng.standard_normal(n))
    put("S", rng.random(q * K))
    put("theta", rng.gamma(hyper["zeta"] + q))
    put("mu", rng.standard_normal())
    put("lambda", rng.random(R))
    # Delta / M / pi need parameters of the NEW state; they are filled lazily by the caller
    return inj, lay


def run_chain(X, y, R, iters, seed, hyper=None, K=64, literal=False, record=("gamma", "xi", "tau2", "mu", "theta", "Delta")):
    """Prior init + `iters` sweeps.  Returns dict of traces (rows = iters+1, row 0 = init)."""
    hyper = dict(O.DEFAULT_HYPER, **(hyper or {}))
    rng = np.random.default_rng(seed)
    n, q = X.shape
    V = int((-1 + math.sqrt(1 + 8 * q)) / 2)
    il = O.init_layout(V, R)
    inj0 = np.empty(il["_total"])

    def put0(name, vals):
        o, s = il[name]
        inj0[o:o + s] = np.asarray(vals).ravel()

    put0("S", rng.exponential(size=q))
    put0("pi", np.stack([rng.gamma([(r + 1) ** hyper["eta"], 1.0, 1.0]) for r in range(R)]))
    put0("lambda", rng.random(R))
    put0("xi", rng.random(V))
    put0("M", np.concatenate([rng.chisquare([hyper["nu"] - i for i in range(R)]),
                              rng.standard_normal(R * (R - 1) // 2)]))
    put0("u", rng.standard_normal(V * R))
    put0("gamma", rng.standard_normal(q))
    st = O.initialize_state(V, R, hyper, inj0)
    tr = {k: [np.array(st[k], dtype=float).copy()] for k in record}
    for _ in range(iters):
        st = sweep_with_rng(st, X, y, V, R, hyper, rng, K, literal)
        for k in record:
            tr[k].append(np.array(st[k], dtype=float).copy())
    return {k: np.stack(v) for k, v in tr.items()}, st


def sweep_with_rng(state, X, y, V, R, hyper, rng, K=64, literal=False):
    """One sweep; the gamma-variate shapes that depend on freshly drawn values are generated in order."""
    n = X.shape[0]
    q = V * (V + 1) // 2
    new = dict(state)
    t = O.update_tau2(X, y, V, state["mu"], state["gamma"], state["u"], state["lam"], state["S"],
                      rng.gamma(n / 2.0 + V * (V + 1) / 4.0))
    new["tau2"] = t["tau2"]
    ux = O.update_u_xi(V, new["tau2"], state["u"], state["lam"], state["S"], state["gamma"], state["Delta"],
                       state["M"], rng.random(V), rng.standard_normal((V, R)), literal)
    new["u"], new["xi"] = ux["u"], ux["xi"]
    g = O.update_gamma(X, y, new["tau2"], new["u"], state["lam"], state["S"], state["mu"],
                       rng.standard_normal(q), rng.standard_normal(n))
    new["gamma"] = g["gamma"]
    d = O.update_D(new["gamma"], new["u"], state["lam"], new["tau2"], state["theta"], rng.random((q, K)))
    new["S"] = d["S"]
    new["theta"] = O.update_theta(new["S"], hyper["zeta"], hyper["iota"], V, rng.gamma(hyper["zeta"] + q))["theta"]
    a = hyper["a_delta"] + new["xi"].sum()
    b = hyper["b_delta"] + V - new["xi"].sum()
    new["Delta"] = O.update_Delta(new["xi"], hyper["a_delta"], hyper["b_delta"],
                                  rng.gamma(a) if a > 0 else 0.0, rng.gamma(b) if b > 0 else 0.0, rng.random())["Delta"]
    df = hyper["nu"] + int(np.sum(np.abs(new["xi"]) > 0.1))
    new["M"] = O.update_M(new["u"], new["xi"], hyper["nu"], rng.chisquare([df - i for i in range(R)]),
                          rng.standard_normal(R * (R - 1) // 2))["M"]
    new["mu"] = O.update_mu(X @ new["gamma"], y, new["tau2"], rng.standard_normal())["mu"]
    new["lam"] = O.update_lambda(new["gamma"], new["u"], new["S"], new["tau2"], state["lam"], state["pi"],
                                 rng.random(R))["lam"]
    al = O.update_pi(new["lam"], hyper["eta"], np.ones((R, 3)))["alpha"]
    new["pi"] = O.update_pi(new["lam"], hyper["eta"], rng.gamma(al))["pi"]
    return new
