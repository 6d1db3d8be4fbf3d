import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from __graft_entry__ import load_package
from oracle import bnr_oracle as O
from test_gpu_parity import make_problem, random_state, sweep_injection
bnr = load_package()
for (V, n) in ((16, 129), (16, 136), (24, 130)):
    R, C, K = 3, 2, 64
    q = V * (V + 1) // 2
    X, y = make_problem(5, V, R, n)
    rng = np.random.default_rng(8)
    states = [random_state(rng, V, R) for _ in range(C)]
    with bnr.Engine(X, y, R, num_chains=C, seed=5, gig_inject_len=K, gamma_mode="nform") as eng:
        eng.enable_aux(True)
        for c, st in enumerate(states):
            eng.set_state_dict(c, st)
        eng.step("gamma")
        for c, st in enumerate(states):
            G = eng.get_aux(c, "G").reshape(n, n).T
            want = (X * st["S"][None, :]) @ X.T + np.eye(n)
            d = np.abs(G - want)
            print(V, n, c, "G err", d.max(), "row128 err", d[128].max(), "rows<128 err", d[:128].max(), "status", eng.status()[c])
