#!/usr/bin/env python
"""Benchmark of the Gibbs hot path (BASELINE.json metric: Gibbs iterations/s summed over all chains).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine (libbnr.so through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's algorithm on the host CPU cores

A "step" is one Gibbs sweep (all ten full conditionals) of EVERY chain held by a GPU.  Workload at any N:
BASELINE config 3 -- synthetic networks with V=100 nodes (q = V(V+1)/2 = 5050 edge coefficients, the reference's
HEAD convention), n=1000 samples, R=7, 64 chains per GPU (weak scaling: N GPUs advance 64*N independent chains,
global chain ids key the RNG; the only exchange is an NCCL all-gather of split-half moments for R-hat).
Inputs are larger than L2 by construction (the per-sweep working set is 64 x 8 MB Gram matrices = 512 MB >> 126 MB),
so no explicit L2 flush is needed between timed iterations.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: V, n, R, chains per GPU
    "c2": dict(V=30, n=500, R=7, chains=16),
    "c3": dict(V=100, n=1000, R=7, chains=64),
    "c4": dict(V=200, n=400, R=9, chains=8),
    "c5": dict(V=50, n=500, R=5, chains=128),
}
CFG_ID = {"c2": 2, "c3": 3, "c4": 4, "c5": 5}


def synth(cfg_name, dense=False):
    """SURVEY 8(d) generator: sparse weighted networks with a low-rank-ish truth; seeded, identical on every rank."""
    cfg = CONFIGS[cfg_name]
    V, n = cfg["V"], cfg["n"]
    rng = np.random.Generator(np.random.Philox(key=20241000 + CFG_ID[cfg_name]))
    q = V * (V + 1) // 2
    xi_true = rng.random(V) < 2.0 / 3.0
    il, ik = np.tril_indices(V)
    # column-major lower triangle incl. diagonal: column k, rows l = k..V-1
    order = np.lexsort((il, ik))
    il, ik = il[order], ik[order]
    B = np.where(xi_true[il] & xi_true[ik] & (il != ik), rng.normal(1.5, 0.9, size=q), 0.0)
    if dense:
        X = rng.normal(size=(n, q))
    else:
        present = rng.random((n, V)) < 0.73
        both = present[:, il] & present[:, ik] & (il != ik)[None, :]
        on = rng.random((n, q)) < 0.9
        X = np.where(both & on, 0.13 + rng.gamma(1.2, 0.2, size=(n, q)), 0.0)
    y = 55.0 + X @ B + rng.normal(0.0, 10.0, size=n)
    return np.asfortranarray(X), y, dict(V=V, n=n, q=q, R=cfg["R"])


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([t.strip() for t in out.strip().split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=mx or None, reasons=sorted(reasons),
                    samples=len(sm))


def fp64_peak_tflops(torch, dev):
    """Measured FP64 GEMM rate (cuBLAS DGEMM through torch.matmul, best of 5 at 4096^3 and 8192x8192x4096):
    the roofline denominator for the DMMA kernels; MEASURED_PEAKS.json only holds HBM and bf16 numbers."""
    best = 0.0
    for (m, n, k) in ((4096, 4096, 4096), (8192, 8192, 4096)):
        a = torch.randn(m, k, dtype=torch.float64, device=dev)
        b = torch.randn(k, n, dtype=torch.float64, device=dev)
        for _ in range(2):
            (a @ b)
        torch.cuda.synchronize(dev)
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize(dev)
            best = max(best, 2.0 * m * n * k / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        del a, b, c
    torch.cuda.empty_cache()
    return best


def run_reference(args, rank):
    """The reference's algorithm on the host cores (oracle/cpu_baseline.py: dense formulation, one chain per process)."""
    if rank != 0:
        return
    from oracle import cpu_baseline as B  # noqa
    X, y, dims = synth(args.config, args.dense)
    chains = CONFIGS[args.config]["chains"]
    sweeps = max(1, min(args.steps, args.ref_sweeps))
    t_all = []
    for _ in range(1):
        its, nproc, wall = B.time_port(np.ascontiguousarray(X), y, dims["R"], chains, sweeps, warm=max(1, min(args.warmup, 1)))
        t_all.append(its)
    val = float(np.mean(t_all))
    line = {
        "impl": "reference", "metric": "gibbs_iters_per_sec_all_chains", "value": val, "unit": "chain-iterations/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * nproc / val,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, dims, chains),
        "cpu_baseline": {"value": val, "unit": "chain-iterations/s", "cores": nproc, "kind": "port",
                         "sample": "%d chains x %d sweeps (+1 warm-up) of the same workload, one chain per process, "
                                   "BLAS threads = 1; Julia absent, NumPy/OpenBLAS port of gibbs_sample!" % (nproc, sweeps)},
        "e2e": {"value": val, "unit": "chain-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def _finite(o):
    """JSON has no NaN / Infinity: non-finite floats become null."""
    if isinstance(o, float):
        return o if math.isfinite(o) else None
    if isinstance(o, dict):
        return {k: _finite(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_finite(v) for v in o]
    return o


def workload_config(args, dims, chains):
    return {"workload": "BASELINE config 3-style synthetic network regression" if args.config == "c3" else args.config,
            "name": args.config, "V": dims["V"], "q": dims["q"], "n": dims["n"], "R": dims["R"],
            "chains_per_gpu": chains, "q_convention": "V(V+1)/2 (reference HEAD, diagonal included)",
            "X": "dense N(0,1)" if args.dense else "sparse weighted networks (SURVEY 8d)",
            "l2": "inputs larger than L2 (per-sweep working set %d MB)" % (chains * (math.ceil(dims["n"] / 128) * 128) ** 2 * 8 // 2 ** 20),
            "parallelism": "chains sharded over GPUs, %d per GPU" % chains}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=6)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--chains", type=int, default=0, help="chains per GPU (default: the config's)")
    ap.add_argument("--total-chains", type=int, default=0,
                    help="strong scaling: this many chains in total, split evenly over the GPUs (BASELINE config 3 "
                         "quotes 64 chains across 1/2/4/8 GPUs); default is weak scaling with the config's chains per GPU")
    ap.add_argument("--chain-groups", type=int, default=0, help="independent stream/graph groups per GPU (0 = library default)")
    ap.add_argument("--gamma-mode", default="auto", choices=["auto", "nform", "qform"])
    ap.add_argument("--dense", action="store_true", help="dense Gaussian X instead of sparse networks")
    ap.add_argument("--ref-sweeps", type=int, default=12, help="bounded CPU sample: sweeps per chain")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-sweeps", type=int, default=5)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package
    bnr = load_package()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))

    X, y, dims = synth(args.config, args.dense)
    chains = args.chains or CONFIGS[args.config]["chains"]
    scaling = "weak"
    if args.total_chains:
        if args.total_chains % world:
            raise SystemExit("--total-chains must be a multiple of the GPU count")
        chains, scaling = args.total_chains // world, "strong"
    V, q, n, R = dims["V"], dims["q"], dims["n"], dims["R"]
    K, Wm = args.steps, max(args.warmup, 3)

    peak = fp64_peak_tflops(torch, dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident timing: `value` ----------------
    rows = Wm + K + 1 + args.profile_sweeps + 2
    eng = bnr.Engine(X, y, R, num_chains=chains, seed=20241018, chain_offset=rank * chains, device=local_rank,
                     trace_rows=rows, trace_full_chains=1, trace_gamma_xi_all=True, chain_groups=args.chain_groups,
                     gamma_mode=args.gamma_mode)
    eng.init_state()
    eng.run(Wm)
    eng.set_moment_window(Wm + 1, K)
    launches0 = eng.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    eng.run(K, sync=True)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    dev_ms = eng.last_run_ms()
    clocks = sampler.stop()
    launches = eng.launch_count() - launches0
    total_ms = max_over_ranks(dev_ms)          # collective: every rank calls it exactly once, here
    step_ms = total_ms / K
    value = world * chains * K / (total_ms * 1e-3)

    # R-hat over all chains of all ranks: NCCL all-gather of split-half moments, reduced identically on every rank
    if K // 2 < 2:
        rx, rg = np.full(V, np.nan), np.full(q, np.nan)        # too few timed draws for a split R-hat
    elif world > 1:
        ptr, cnt = eng.moments_device()
        mine = torch.empty(cnt, dtype=torch.float64, device=dev)
        eng.export_moments(mine.data_ptr())
        allm = torch.empty(cnt * world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allm, mine)
        torch.cuda.synchronize(dev)
        rx, rg = eng.rhat_from_moments(allm.data_ptr(), chains * world, K // 2)
    else:
        rx, rg = eng.rhat()
    status = eng.status()

    # gamma ESS/s of EVERY edge coefficient, on the device (bnr_ess_*: per-chain-centred autocovariances by direct
    # lag sums + Geyer's initial monotone sequence); across GPUs the per-rank autocovariance sums and chain means are
    # all-gathered over NCCL and reduced identically on every rank
    max_lag = min(255, K - 1)
    ess_x = ess_g = None
    lag = 0
    if K >= 4:
        eng.ess_accumulate(Wm + 1, K, max_lag)
        (pa, na), (pm, nm), lag = eng.ess_device()
    if K < 4:
        ess_g = np.full(q, np.nan)
    elif world > 1:
        a_mine = torch.empty(na, dtype=torch.float64, device=dev)
        m_mine = torch.empty(nm, dtype=torch.float64, device=dev)
        eng.export_ess(a_mine.data_ptr(), m_mine.data_ptr())
        a_all = torch.empty(na * world, dtype=torch.float64, device=dev)
        m_all = torch.empty(nm * world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(a_all, a_mine)
        dist.all_gather_into_tensor(m_all, m_mine)
        torch.cuda.synchronize(dev)
        ess_x, ess_g = eng.ess_from_stats(a_all.data_ptr(), world, m_all.data_ptr(), chains * world, K, lag)
    else:
        ess_x, ess_g = eng.ess_from_stats(pa, 1, pm, chains, K, lag)
    ess_med = float(np.nanmedian(ess_g)) if np.isfinite(ess_g).any() else float("nan")
    ess_min = float(np.nanmin(ess_g)) if np.isfinite(ess_g).any() else float("nan")

    # ---------------- per-phase CUDA-event profile of eager sweeps (roofline numerator) ----------------
    phases = {}
    for _ in range(max(1, args.profile_sweeps)):
        for k, v in eng.profile_sweep().items():
            phases.setdefault(k, []).append(v)
    phases = {k: float(np.mean(v)) for k, v in phases.items()}
    gmode = eng.gamma_mode
    if gmode == "nform":
        # dominant kernel: G = X D X' + I.  SURVEY 8(d): n^2 q flops per chain-iteration (lower half, mul+add)
        dom_kernel = "k_gram_syrk (X diag(S) X' + I, DMMA m8n8k4, TMA ring)"
        dom_ms = phases["syrk"]
        dom_flops = chains * float(n) * n * q
    else:
        # q-form: the batched blocked Cholesky of the q x q precision dominates; q^3/3 flops per chain-iteration
        dom_kernel = "blocked Cholesky of P = (X'X + D^-1)/tau2 (k_chol_update + k_potf2_inv + k_trsm_dmma)"
        dom_ms = phases["cholesky"]
        dom_flops = chains * float(q) ** 3 / 3.0
    achieved = dom_flops / (dom_ms * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "syrk_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(args.config)
        except Exception:
            traffic = None
    eng.close()

    # ---------------- end to end through the public API with HOST buffers: `e2e` ----------------
    # The call a user makes: Fit(X, y, R; nburn, nsamples, num_chains, seed) -> Results, then Summary(Results).
    # Inside the timed region: handle creation, H2D of X and y, prior init, K sweeps of every chain, streamed R-hat,
    # Summary reduced on the device, D2H of chain 1's gamma / xi table, the R-hat vectors and the Summary statistics.
    nsamp_e = min(max(2, K // 2), K)
    nburn_e = K + 1 - nsamp_e                      # nburn + nsamples rows = prior row + K sweeps
    barrier()
    t0 = time.perf_counter()
    res = bnr.Fit(X, y, R, nburn=nburn_e, nsamples=nsamp_e, num_chains=chains, seed=7, x_transform=False,
                  filename=None, psrf_cutoff=float("inf"), device=local_rank, chain_offset=rank * chains,
                  return_state="gamma_xi")
    summ = bnr.Summary(res) if nsamp_e >= 40 else None     # the reference's Summary needs >= 20 draws per tail index
    torch.cuda.synchronize(dev)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_val = world * chains * K / e2e_s
    h2d = (X.nbytes + y.nbytes) / K
    d2h = (res.state["gamma"].nbytes + res.state["xi"].nbytes + 8 * (V + q) + 8 * (3 * q + V)) / K
    assert (summ is None or len(summ.edge_coef["estimate"]) == q) and res.extra["tot_generated"] == K + 1

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import cpu_baseline as B  # noqa
        sweeps = max(2, min(args.ref_sweeps, 8))
        its, nproc, wall = B.time_port(np.ascontiguousarray(X), y, R, chains, sweeps, warm=1)
        cpu = {"value": its, "unit": "chain-iterations/s", "cores": nproc, "kind": "port",
               "sample": "%d chains x %d sweeps (+1 warm-up), one chain per process, BLAS threads = 1, %.1f s wall; "
                         "NumPy/OpenBLAS port of the reference's dense formulation (Julia absent)" % (nproc, sweeps, wall)}

    if rank == 0:
        algo_flops = n * n * q + n ** 3 / 3.0 + 2.0 * n * n + 10.0 * n * q   # n-form, SURVEY 8(d)
        line = {
            "metric": "gibbs_iters_per_sec_all_chains", "value": value, "unit": "chain-iterations/s",
            "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, dims, chains),
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "chain-iterations/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "includes": "Fit(X, y, R; ...) + Summary through the public API: handle creation "
                    "(device buffers come from libbnr's cache, warmed by the device-resident leg above), H2D of X,y, "
                    "prior init, K sweeps, streamed R-hat, device Summary, D2H of chain-1 gamma/xi table, handle teardown"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": dom_kernel,
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": traffic, "ms_per_launch": dom_ms, "flops_per_launch": dom_flops,
                         "peak_source": "cuBLAS DGEMM measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                         "peak_theoretical": 128 * 148 * 1.965e9 / 1e12,
                         "frac_of_theoretical": achieved / (128 * 148 * 1.965e9 / 1e12),
                         "note": "theoretical = 128 FP64 flop/clk/SM (DMMA and DFMA alike) x 148 SMs x 1.965 GHz"},
            "cpu_baseline": cpu,
            "gamma_ess_per_sec": {"median": ess_med / (total_ms * 1e-3),
                                  "min": ess_min / (total_ms * 1e-3),
                                  "edges": int(q), "draws_per_chain": K, "chains": chains * world, "max_lag": lag,
                                  "note": "device-side multi-chain Geyer ESS of every gamma_j over the timed draws "
                                          "(short window right after warm-up: indicative, not a converged-run figure)"},
            "gamma_mode": gmode,
            "algorithmic_tflops": value * algo_flops / 1e12,
            "phases_ms": phases,
            "wall_ms_per_step": wall_ms / K,
            "rhat": {"max_gamma": float(np.nanmax(rg)) if np.isfinite(rg).any() else None, "max_xi": float(np.nanmax(rx[np.isfinite(rx)])) if np.isfinite(rx).any() else None,
                     "chains": chains * world},
            "status_or": int(np.bitwise_or.reduce(status)),
        }
        print(json.dumps(_finite(line)), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
