"""Long run at BASELINE config 3 (or --config): chain health, convergence and recovery of the simulation truth."""
import sys, time, numpy as np
sys.path.insert(0, '.')
import bench
from __graft_entry__ import load_package
bnr = load_package()
cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
nburn, nsamp = int(sys.argv[2]) if len(sys.argv) > 2 else 1500, int(sys.argv[3]) if len(sys.argv) > 3 else 1000
X, y, dims = bench.synth(cfg)
chains = bench.CONFIGS[cfg]["chains"]
# truth (same generator stream as bench.synth)
V = dims["V"]; q = dims["q"]
rng = np.random.Generator(np.random.Philox(key=20241000 + bench.CFG_ID[cfg]))
xi_true = rng.random(V) < 2.0 / 3.0
t0 = time.perf_counter()
res = bnr.Fit(X, y, dims["R"], nburn=nburn, nsamples=nsamp, num_chains=chains, seed=11, x_transform=False,
              filename=None, psrf_cutoff=1e9, return_state="none")
dt = time.perf_counter() - t0
out = bnr.Summary(res)
prob = out.prob_nodes["probability"]
st = res.extra["status"]
print(cfg, "chains", chains, "sweeps", nburn + nsamp, "wall s %.1f" % dt, "it/s %.0f" % (chains * (nburn + nsamp) / dt))
print("status bits OR", int(np.bitwise_or.reduce(st)), "streamed", res.extra["rhat_streamed"], "gamma_mode", res.extra["gamma_mode"])
rg = np.asarray(res.rhatγ.γ); rx = np.asarray(res.rhatξ.ξ)
print("rhat gamma max %.3f median %.3f; rhat xi finite max %.3f" % (np.nanmax(rg), np.nanmedian(rg), np.nanmax(rx[np.isfinite(rx)]) if np.isfinite(rx).any() else float('nan')))
print("node calls agree with truth: %d / %d" % (int(((prob > 0.5) == xi_true).sum()), V))
sig = (out.edge_coef["lower_bound"] > 0) | (out.edge_coef["upper_bound"] < 0)
print("significant edges", int(sig.sum()), "of", q)
