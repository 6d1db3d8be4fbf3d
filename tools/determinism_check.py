#!/usr/bin/env python
"""Development check: the same seed must give bit-identical chains run after run and for any chain grouping."""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from __graft_entry__ import load_package


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c3")
    ap.add_argument("--chains", type=int, default=0)
    ap.add_argument("--sweeps", type=int, default=100)
    ap.add_argument("--repeats", type=int, default=3)
    ap.add_argument("--groups", default="0,0,1")
    args = ap.parse_args()
    bnr = load_package()
    X, y, dims = bench.synth(args.config)
    chains = args.chains or bench.CONFIGS[args.config]["chains"]
    ref = None
    for g in [int(v) for v in args.groups.split(",")]:
        with bnr.Engine(X, y, dims["R"], num_chains=chains, seed=11, device=0, trace_rows=0, chain_groups=g) as eng:
            eng.init_state()
            eng.run(args.sweeps, sync=True)
            st = [np.concatenate([np.ravel(eng.get_state(c, k)) for k in ("gamma", "S", "u", "tau2", "M", "lam")]) for c in range(chains)]
            st = np.stack(st)
            status = eng.status()
            if ref is None:
                ref = st
                print("groups", eng.chain_groups, "reference run; status_or", int(np.bitwise_or.reduce(status)))
            else:
                bad = [c for c in range(chains) if not np.array_equal(ref[c], st[c])]
                print("groups", eng.chain_groups, "chains that differ:", bad[:16], "of", chains,
                      "max rel diff %.3e" % (np.max(np.abs(ref - st)) / np.max(np.abs(ref))))


if __name__ == "__main__":
    main()
