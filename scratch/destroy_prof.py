import sys, time, numpy as np
sys.path.insert(0, '.')
import bench
from __graft_entry__ import load_package
bnr = load_package()
X, y, dims = bench.synth("c3")
for C, rows in ((64, 101), (64, 0), (8, 101), (64, 101)):
    t = time.perf_counter()
    eng = bnr.Engine(X, y, 7, num_chains=C, seed=7, trace_rows=rows, trace_full_chains=0, trace_gamma_xi_all=False, trace_gamma_xi_chains=1)
    t1 = time.perf_counter()
    eng.init_state(); eng.run(4)
    t2 = time.perf_counter()
    eng.close()
    t3 = time.perf_counter()
    print("C=%d rows=%d create %.1f ms  close %.1f ms" % (C, rows, (t1 - t) * 1e3, (t3 - t2) * 1e3))
