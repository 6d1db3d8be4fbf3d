import sys, time, numpy as np
sys.path.insert(0, '.')
import bench
from __graft_entry__ import load_package
bnr = load_package()
X, y, dims = bench.synth("c3")
R = dims["R"]
def T(msg, t0):
    print("%-28s %.1f ms" % (msg, (time.perf_counter() - t0) * 1e3)); return time.perf_counter()
for rep in range(2):
    print("--- rep", rep)
    t = time.perf_counter()
    eng = bnr.Engine(X, y, R, num_chains=64, seed=7, trace_rows=101, trace_full_chains=0, trace_gamma_xi_all=False, trace_gamma_xi_chains=1)
    t = T("Engine()", t)
    eng.init_state(); t = T("init_state", t)
    eng.set_moment_window(51, 50); t = T("set_moment_window", t)
    eng.trace_row = 1; t = T("trace_row", t)
    eng.run(100); t = T("run(100)", t)
    rx, rg = eng.rhat(); t = T("rhat", t)
    s = eng.summary(0, 51, 50, 1, 49); t = T("summary", t)
    g = eng.get_trace(0, "gamma", 0, 101); x = eng.get_trace(0, "xi", 0, 101); t = T("get_trace", t)
    st = eng.status(); eng.close(); t = T("status+close", t)
import cProfile, pstats
for rep in range(2):
    t = time.perf_counter()
    res = bnr.Fit(X, y, R, nburn=51, nsamples=50, num_chains=64, seed=7, x_transform=False, filename=None,
                  psrf_cutoff=float("inf"), return_state="gamma_xi")
    out = bnr.Summary(res)
    print("Fit+Summary %.1f ms" % ((time.perf_counter() - t) * 1e3))
pr = cProfile.Profile(); pr.enable()
res = bnr.Fit(X, y, R, nburn=51, nsamples=50, num_chains=64, seed=7, x_transform=False, filename=None, psrf_cutoff=float("inf"), return_state="gamma_xi")
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
