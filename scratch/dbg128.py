import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from __graft_entry__ import load_package
from oracle import bnr_oracle as O
from test_gpu_parity import make_problem, random_state, sweep_injection
bnr = load_package()
for n in (128, 129, 135, 256, 384):
    V, R, C, K = 6, 3, 2, 64
    q = V * (V + 1) // 2
    X, y = make_problem(5, V, R, n)
    rng = np.random.default_rng(8)
    states = [random_state(rng, V, R) for _ in range(C)]
    injs = [sweep_injection(rng, n, V, R, K) for _ in range(C)]
    inj, lay = np.stack([i[0] for i in injs]), injs[0][1]
    with bnr.Engine(X, y, R, num_chains=C, seed=5, gig_inject_len=K, gamma_mode="nform") as eng:
        eng.enable_aux(True)
        for c, st in enumerate(states):
            eng.set_state_dict(c, st)
        eng.set_injection(inj)
        eng.step("gamma")
        for c, st in enumerate(states):
            o1, s1 = lay["gamma_z1"]; o2, s2 = lay["gamma_z2"]
            want = O.update_gamma(X, y, st["tau2"], st["u"], st["lam"], st["S"], st["mu"], inj[c, o1:o1+s1], inj[c, o2:o2+s2])
            G = eng.get_aux(c, "G").reshape(n, n).T
            L = eng.get_aux(c, "G_chol").reshape(n, n).T
            a4 = eng.get_aux(c, "a4")
            Lw = np.linalg.cholesky(want["G"])
            print(n, c, "G", np.abs(G - want["G"]).max(), "L", np.abs(L - Lw).max(), "a4", np.abs(a4 - want["a4"]).max(),
                  "gamma", np.abs(eng.get_state(c, "gamma")[:, 0] - want["gamma"]).max(), "status", eng.status()[c])
            if np.abs(L - Lw).max() > 1e-6:
                bad = np.argwhere(np.abs(L - Lw) > 1e-6)
                print("   bad L entries rows", bad[:, 0].min(), bad[:, 0].max(), "cols", bad[:, 1].min(), bad[:, 1].max(), len(bad))
print("---- detail n=129")
n = 129
V, R, C, K = 6, 3, 1, 64
X, y = make_problem(5, V, R, n)
rng = np.random.default_rng(8)
st = random_state(rng, V, R)
inj, lay = sweep_injection(rng, n, V, R, K)
with bnr.Engine(X, y, R, num_chains=1, seed=5, gig_inject_len=K, gamma_mode="nform") as eng:
    eng.enable_aux(True)
    eng.set_state_dict(0, st)
    eng.set_injection(inj[None, :])
    eng.step("gamma")
    o1, s1 = lay["gamma_z1"]; o2, s2 = lay["gamma_z2"]
    want = O.update_gamma(X, y, st["tau2"], st["u"], st["lam"], st["S"], st["mu"], inj[o1:o1+s1], inj[o2:o2+s2])
    L = eng.get_aux(0, "G_chol").reshape(n, n).T
    Lw = np.linalg.cholesky(want["G"])
    d = np.abs(L - Lw)
    print("rows 0..127 max err", np.nanmax(d[:128]), "nan count", np.isnan(L).sum())
    print("row 128 first 6:", L[128, :6], "want", Lw[128, :6])
    print("row 128 last 4:", L[128, 125:129], "want", Lw[128, 125:129])
