"""Level-1 parity, the corners: BASELINE-size injected sweeps, the degenerate / failure arms of the samplers, and the
tolerance rule for quantities behind the n x n (q x q) solve, measured against a long-double arbiter."""
import json
import math
import os

import numpy as np
import pytest

from oracle import bnr_oracle as O
from oracle import extended as E
from parity_util import EPS, solve_tol, note as _note
from test_gpu_parity import make_problem, random_state, sweep_injection, _init_injection, close

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------------------------------------------------------
# BASELINE shapes: one injected sweep of two chains against the oracle (configs 3, 4, 5 in the n-form, 2 in the q-form)
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,V,R,n,mode", [("c3", 100, 7, 1000, "nform"), ("c4", 200, 9, 400, "nform"),
                                             ("c5", 50, 5, 500, "nform"), ("c2", 30, 7, 500, "qform")])
def test_baseline_size_sweep_injected(bnr, name, V, R, n, mode):
    C, K = 2, 48
    X, y = make_problem(V + n, V, R, n, dense=False)
    rng = np.random.default_rng(V)
    hyper = dict(O.DEFAULT_HYPER)
    with bnr.Engine(X, y, R, num_chains=C, seed=1, gig_inject_len=K, gamma_mode=mode) as eng:
        assert eng.gamma_mode == mode
        init = np.stack([_init_injection(rng, V, R) for _ in range(C)])
        eng.set_injection(init)
        eng.init_state()
        inj = np.stack([sweep_injection(rng, n, V, R, K)[0] for _ in range(C)])
        eng.set_injection(inj)
        eng.run(1)
        for c in range(C):
            st = O.initialize_state(V, R, hyper, init[c])
            new, aux = O.gibbs_sweep(st, X, y, V, R, hyper, inj[c], K, literal=False,
                                     gamma_form="q" if mode == "qform" else "n")
            cond = np.linalg.cond(aux["gamma"]["P" if mode == "qform" else "G"])
            tol = solve_tol(cond)
            got = eng.get_state_dict(c)
            for k in ("tau2", "xi", "lam"):
                close(got[k], new[k], rtol=1e-10, msg="%s %s" % (name, k))
            worst = 0.0
            for k in ("u", "gamma", "S", "theta", "Delta", "M", "mu", "pi"):
                a, b = np.asarray(got[k], dtype=float), np.asarray(new[k], dtype=float)
                scale = np.max(np.abs(b)) if b.size else 1.0
                err = float(np.max(np.abs(a - b)) / scale) if b.size else 0.0
                worst = max(worst, err)
                np.testing.assert_allclose(a, b, rtol=tol, atol=tol * scale, err_msg="%s %s" % (name, k))
            _note(test="baseline_sweep", config=name, chain=c, cond=cond, worst_rel_err=worst, c=worst / (cond * EPS))
        assert not eng.status().any()


# ------------------------------------------------------------------------------------------------------------------
# the tolerance rule: GPU and float64 oracle against the long-double arbiter, for growing condition numbers
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.skipif(np.finfo(np.longdouble).eps > 1e-18, reason="no 80-bit long double on this platform")
@pytest.mark.parametrize("mode", ["nform", "qform"])
@pytest.mark.parametrize("s_scale", [1.0, 1e3, 1e6])
def test_solve_error_follows_cond_times_eps(bnr, mode, s_scale):
    V, R, n, C = 12, 4, 150, 2
    X, y = make_problem(77, V, R, n)
    rng = np.random.default_rng(int(math.log10(s_scale)) + 5)
    K = 8
    with bnr.Engine(X, y, R, num_chains=C, seed=2, gig_inject_len=K, gamma_mode=mode) as eng:
        eng.enable_aux(True)
        states = []
        for c in range(C):
            st = random_state(rng, V, R)
            st["S"] = st["S"] * s_scale
            states.append(st)
            eng.set_state_dict(c, st)
        inj = np.stack([sweep_injection(rng, n, V, R, K)[0] for _ in range(C)])
        eng.set_injection(inj)
        lay = O.draw_layout(n, V, R, K)
        eng.step("gamma")
        for c, st in enumerate(states):
            z1 = inj[c, lay["gamma_z1"][0]:lay["gamma_z1"][0] + lay["gamma_z1"][1]]
            z2 = inj[c, lay["gamma_z2"][0]:lay["gamma_z2"][0] + lay["gamma_z2"][1]]
            if mode == "nform":
                ld = E.update_gamma_ld(X, y, st["tau2"], st["u"], st["lam"], st["S"], st["mu"], z1, z2)
                f64 = O.update_gamma(X, y, st["tau2"], st["u"], st["lam"], st["S"], st["mu"], z1, z2)
                cond = np.linalg.cond(f64["G"])
            else:
                ld = E.update_gamma_qform_ld(X, y, st["tau2"], st["u"], st["lam"], st["S"], st["mu"], z1)
                f64 = O.update_gamma_qform(X, y, st["tau2"], st["u"], st["lam"], st["S"], st["mu"], z1)
                cond = np.linalg.cond(f64["P"])
            got = eng.get_state(c, "gamma")[:, 0]
            e_gpu, e_f64 = E.rel_err(got, ld["gamma"]), E.rel_err(f64["gamma"], ld["gamma"])
            m = n if mode == "nform" else V * (V + 1) // 2
            Lg = eng.get_aux(c, "G_chol").reshape(m, m).T
            e_chol = E.rel_err(Lg, ld["L"])
            _note(test="solve_error", mode=mode, s_scale=s_scale, cond=cond, e_gpu=e_gpu, e_f64=e_f64, e_chol=e_chol,
                  c_gpu=e_gpu / (cond * EPS), c_f64=e_f64 / (cond * EPS), c_chol=e_chol / (cond * EPS))
            assert e_gpu <= solve_tol(cond), (e_gpu, cond)
            assert e_chol <= solve_tol(cond), (e_chol, cond)
            assert e_f64 <= solve_tol(cond)


# ------------------------------------------------------------------------------------------------------------------
# degenerate and failure arms
# ------------------------------------------------------------------------------------------------------------------
def test_jitter_ladder_matches_reference(bnr):
    """src/gibbs.jl:322-347 through the device routine the sweep uses (bnr_test_chol_jitter)."""
    X, y = make_problem(1, 5, 3, 20)
    rng = np.random.default_rng(0)
    with bnr.Engine(X, y, 3, num_chains=1, seed=1) as eng:
        for R in (3, 7, 16):
            Q, _ = np.linalg.qr(rng.normal(size=(R, R)))
            for lam_min, want_jit, fails in ((0.5, 0, False), (-5e-6, 1, False), (-3e-5, 2, False), (-1e-3, 2, True)):
                ev = np.concatenate([[lam_min], rng.uniform(0.5, 2.0, size=R - 1)])
                A = (Q * ev[None, :]) @ Q.T
                A = 0.5 * (A + A.T)
                used, L, st = eng.test_chol_jitter(A)
                if fails:
                    with pytest.raises(np.linalg.LinAlgError):
                        O.chol_with_jitter(A)
                    assert st & 2                                     # BNR_ST_SIGMA_NOTPD: the reference throws here
                    continue
                Lw, Aw, jit = O.chol_with_jitter(A)
                assert jit == want_jit
                assert bool(st & 1) == (jit > 0) and not (st & 2)
                np.testing.assert_allclose(used, Aw, rtol=0, atol=1e-15)
                np.testing.assert_allclose(L, Lw, rtol=1e-9, atol=1e-12)


def test_nan_weight_falls_back_to_a_fair_coin(bnr):
    """update_xi (src/gibbs.jl:385-402): a NaN mixture weight draws xi ~ Bernoulli(0.5) from the node's uniform."""
    V, R, n, C, K = 6, 3, 20, 2, 8
    X, y = make_problem(4, V, R, n)
    rng = np.random.default_rng(3)
    with bnr.Engine(X, y, R, num_chains=C, seed=5, gig_inject_len=K, gamma_mode="nform") as eng:
        states = [random_state(rng, V, R) for _ in range(C)]
        for st in states:
            st["xi"][:] = 1.0
            st["u"] = rng.normal(size=(R, V))
            st["lam"] = np.ones(R)
            st["gamma"][O.tri_index(3, 1, V)] = np.inf          # edge (3, 1): nodes 1 and 3 get an infinite b vector
        for c, st in enumerate(states):
            eng.set_state_dict(c, st)
        inj = np.stack([sweep_injection(rng, n, V, R, K)[0] for _ in range(C)])
        eng.set_injection(inj)
        eng.step("tau2")      # (tau2 itself becomes inf/NaN; the coin flip below does not depend on it)
        eng.step("u_xi")
        lay = O.draw_layout(n, V, R, K)
        o, s = lay["uxi"]
        for c in range(C):
            ups = inj[c, o:o + s].reshape(V, R + 1)[:, 0]
            xi = eng.get_state(c, "xi")[:, 0]
            for k in (1, 3):
                assert xi[k] == (1.0 if ups[k] <= 0.5 else 0.0)
        assert (eng.status() & 32).all()                            # BNR_ST_NAN


def test_gig_degenerate_arms(bnr):
    """src/gig.jl:15-26: chi < 10 eps -> Gamma(lambda, psi/2) scale convention; psi < 10 eps -> 1 / Gamma(lambda, chi/2)."""
    V, R, n, C, K = 6, 3, 20, 2, 8
    X, y = make_problem(5, V, R, n)
    rng = np.random.default_rng(8)
    with bnr.Engine(X, y, R, num_chains=C, seed=5, gig_inject_len=K, gamma_mode="nform") as eng:
        eng.enable_aux(True)
        states = [random_state(rng, V, R) for _ in range(C)]
        # chain 0: node 2 switched off (u_2 = 0 -> W = 0 on its edges) and gamma = 0 there: chi = 0 exactly
        states[0]["u"][:, 2] = 0.0
        states[0]["xi"][2] = 0.0
        zero_edges = [O.tri_index(max(l, 2), min(l, 2), V) for l in range(V)]
        states[0]["gamma"][zero_edges] = 0.0
        # chain 1: theta below 10 eps
        states[1]["theta"] = 1e-16
        for c, st in enumerate(states):
            eng.set_state_dict(c, st)
        inj = np.stack([sweep_injection(rng, n, V, R, K)[0] for _ in range(C)])
        eng.set_injection(inj)
        eng.step("D")
        lay = O.draw_layout(n, V, R, K)
        o, s = lay["S"]
        q = V * (V + 1) // 2
        for c, st in enumerate(states):
            want = O.update_D(st["gamma"], st["u"], st["lam"], st["tau2"], st["theta"], inj[c, o:o + s].reshape(q, K))
            got = eng.get_state(c, "S")[:, 0]
            np.testing.assert_allclose(got, want["S"], rtol=1e-10)
            np.testing.assert_array_equal(eng.get_aux(c, "gig_used"), want["used"])
            if c == 0:
                assert all(want["branch"][j] == "degenerate_chi" for j in zero_edges)
            else:
                assert set(want["branch"]) == {"degenerate_psi"}
        assert not (eng.status() & (8 | 16)).any()


@pytest.mark.parametrize("a_delta,b_delta,xi_val,want", [(0.0, 1.0, 0.0, 0.0), (1.0, 0.0, 1.0, 1.0)])
def test_sample_beta_degenerate_arms(bnr, a_delta, b_delta, xi_val, want):
    """sample_Beta (src/gibbs.jl:130-140): a = 0 -> Delta = 0, b = 0 -> Delta = 1, no variate consumed."""
    V, R, n, C, K = 6, 3, 20, 2, 8
    X, y = make_problem(6, V, R, n)
    rng = np.random.default_rng(9)
    with bnr.Engine(X, y, R, num_chains=C, seed=5, gig_inject_len=K, a_delta=a_delta, b_delta=b_delta) as eng:
        eng.enable_aux(True)
        for c in range(C):
            st = random_state(rng, V, R)
            st["xi"][:] = xi_val
            eng.set_state_dict(c, st)
        eng.set_injection(np.stack([sweep_injection(rng, n, V, R, K)[0] for _ in range(C)]))
        eng.step("Delta")
        for c in range(C):
            assert eng.get_state(c, "Delta")[0, 0] == want
            a, b = eng.get_aux(c, "delta_params")
            assert (a, b) == (a_delta + V * xi_val, b_delta + V * (1 - xi_val))
            ref = O.update_Delta(np.full(V, xi_val), a_delta, b_delta, 1.0, 1.0)
            assert ref["Delta"] == want


@pytest.mark.parametrize("mode", ["nform", "qform"])
def test_lost_positive_definiteness_is_flagged(bnr, mode):
    """The reference's cholesky(...) throws PosDefException when X D X' + I (src/gibbs.jl:434) is not positive definite;
    the engine cannot throw from a kernel: the panel kernel meets a non-positive pivot and sets BNR_ST_G_NOTPD for that
    chain only (negative scales S make the matrix of chain 0 indefinite; chain 1 stays clean)."""
    V, R, n, C, K = 8, 3, 40, 2, 8
    X, y = make_problem(9, V, R, n)
    rng = np.random.default_rng(12)
    with bnr.Engine(X, y, R, num_chains=C, seed=5, gig_inject_len=K, gamma_mode=mode) as eng:
        states = [random_state(rng, V, R) for _ in range(C)]
        # n-form: G = I - 100 X X';  q-form: P = (X'X - 1000 I) / tau2
        states[0]["S"] = np.full_like(states[0]["S"], -100.0 if mode == "nform" else -1e-3)
        for c, st in enumerate(states):
            eng.set_state_dict(c, st)
        inj = np.stack([sweep_injection(rng, n, V, R, K)[0] for _ in range(C)])
        eng.set_injection(inj)
        eng.step("gamma")
        st = eng.status()
        assert st[0] & 4, st                                   # BNR_ST_G_NOTPD
        assert not (st[1] & 4), st
