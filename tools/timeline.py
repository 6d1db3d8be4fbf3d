#!/usr/bin/env python
"""Kernel timeline of steady-state sweeps (development tool, not part of the product).

CUPTI (through torch.profiler) records every kernel of the process, including the graph-launched kernels of libbnr.so.
Prints, for one window of `--sweeps` sweeps after warm-up: per stream the kernels in start order (offset, duration, grid),
per kernel name the summed duration, and the number of SM-busy CTAs over time is left to the reader.

    python tools/timeline.py --config c3 --chains 8 [--chain-groups 4] [--sweeps 2] [--out gpurun_out/tl.txt]
"""
import argparse
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c3")
    ap.add_argument("--chains", type=int, default=0)
    ap.add_argument("--chain-groups", type=int, default=0)
    ap.add_argument("--gamma-mode", default="auto")
    ap.add_argument("--sweeps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--out", default="")
    args = ap.parse_args()

    import torch
    from torch.profiler import profile, ProfilerActivity
    import bench
    from __graft_entry__ import load_package
    bnr = load_package()
    X, y, dims = bench.synth(args.config)
    chains = args.chains or bench.CONFIGS[args.config]["chains"]
    eng = bnr.Engine(X, y, dims["R"], num_chains=chains, seed=1, device=0, trace_rows=0,
                     chain_groups=args.chain_groups, gamma_mode=args.gamma_mode)
    eng.init_state()
    eng.run(args.warmup, sync=True)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        eng.run(args.sweeps, sync=True)
        torch.cuda.synchronize()
    path = os.path.join(tempfile.mkdtemp(), "trace.json")
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
    ev.sort(key=lambda e: e["ts"])
    out = open(args.out, "w") if args.out else sys.stdout
    if not ev:
        print("no kernels recorded", file=out)
        return
    t0 = ev[0]["ts"]
    tend = max(e["ts"] + e["dur"] for e in ev)
    print(f"# {args.config} chains={chains} groups={eng.chain_groups} sweeps={args.sweeps}: window {tend - t0:.1f} us, "
          f"{(tend - t0) / args.sweeps:.1f} us per sweep, {len(ev)} kernels", file=out)
    streams = sorted({e["args"].get("stream") for e in ev})
    sid = {s: i for i, s in enumerate(streams)}
    tot = {}
    for e in ev:
        nm = e["name"].split("(")[0].replace("bnr::", "").replace("void ", "")
        tot[nm] = tot.get(nm, 0.0) + e["dur"]
    print("# summed durations (us per sweep):", file=out)
    for nm, d in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"#   {nm:32s} {d / args.sweeps:9.1f}", file=out)
    print("# start_us  dur_us  end_us  stream  grid  name", file=out)
    for e in ev:
        nm = e["name"].split("(")[0].replace("bnr::", "").replace("void ", "")
        g = e["args"].get("grid")
        print(f"{e['ts'] - t0:10.1f} {e['dur']:8.1f} {e['ts'] - t0 + e['dur']:10.1f}  s{sid[e['args'].get('stream')]:<3d} {str(g):16s} {nm}",
              file=out)
    eng.close()


if __name__ == "__main__":
    main()
