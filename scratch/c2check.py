"""Level-2 check at BASELINE config 2 (q-form on the GPU): node inclusion probabilities vs the CPU oracle chain."""
import sys, time, numpy as np
sys.path.insert(0, '.')
import bench
from oracle import chain as OC
from __graft_entry__ import load_package
bnr = load_package()
X, y, dims = bench.synth("c2")
X = np.ascontiguousarray(X)
V, R = dims["V"], dims["R"]
rng = np.random.Generator(np.random.Philox(key=20241000 + 2))
xi_true = rng.random(V) < 2.0 / 3.0
res = {}
for mode in ("qform", "nform"):
    with bnr.Engine(X, y, R, num_chains=16, seed=5, trace_rows=10001, trace_full_chains=0, gamma_mode=mode) as eng:
        eng.init_state(); eng.run(10000)
        res[mode] = np.mean([eng.get_trace(c, "xi", 6001, 10001)[:, :, 0].mean(axis=0) for c in range(16)], axis=0)
t0 = time.time()
tr = [OC.run_chain(X, y, R, 2500, seed=s, record=("xi",))[0]["xi"][1001:].reshape(-1, V).mean(axis=0) for s in (1, 2)]
print("oracle time", time.time() - t0)
ora = np.mean(tr, axis=0)
np.set_printoptions(precision=2, suppress=True, linewidth=200)
print("truth ", xi_true.astype(int))
print("qform ", res["qform"])
print("nform ", res["nform"])
print("oracle", ora)
print("max |qform - nform|", np.abs(res["qform"] - res["nform"]).max(), " max |gpu - oracle|", np.abs(res["nform"] - ora).max())
