"""Host side of `Fit!` / `Summary` (src/gibbs.jl:725-751, 897-1020, 1051-1198, 1214-1250).

Only orchestration lives here: keyword handling, parameters.log, X vectorisation (setup_X!), the row
bookkeeping of run! (purge_burn ring), the PSRF-driven "extend burn-in" and "doubling" control loops and the
Summary tables.  Every Gibbs iteration, the traces and the R-hat reduction run on the GPU inside libbnr.
"""
import datetime
import math
import random

import numpy as np

from .engine import Engine

_UNI = {"tau2": "τ²", "u": "u", "xi": "ξ", "gamma": "γ", "S": "S", "theta": "θ", "Delta": "Δ", "M": "M",
        "mu": "μ", "lam": "λ", "pi": "πᵥ"}
_ASCII = {v: k for k, v in _UNI.items()}
_ASCII.update({"lambda": "lam", "pi_v": "pi", "tau²": "tau2"})


class Table(dict):
    """Column container standing in for TypedTables.Table: columns by reference name (state["γ"], state.γ)
    or ASCII alias (state.gamma); each column is a Fortran-ordered (rows, d1, d2) Float64 array."""

    def _key(self, name):
        return _ASCII.get(name, name)

    def __getitem__(self, name):
        return dict.__getitem__(self, self._key(name))

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __len__(self):
        for v in self.values():
            return v.shape[0]
        return 0


class Results:
    """src/gibbs.jl:23-29: state table of the first chain, PSRF tables, burn_in, sampled."""

    def __init__(self, state, rhat_xi, rhat_gamma, burn_in, sampled, extra=None):
        self.state = state
        self.rhatξ = Table(xi=np.asarray(rhat_xi))
        self.rhatγ = Table(gamma=np.asarray(rhat_gamma))
        self.burn_in = int(burn_in)
        self.sampled = int(sampled)
        self.extra = extra or {}

    rhat_xi = property(lambda self: self.rhatξ)
    rhat_gamma = property(lambda self: self.rhatγ)

    def __repr__(self):
        return repr(Summary(self))


class BNRSummary:
    """src/gibbs.jl:39-43.  edge_coef / prob_nodes are dicts of columns (DataFrame stand-ins)."""

    def __init__(self, edge_coef, prob_nodes, ci_level):
        self.edge_coef = edge_coef
        self.prob_nodes = prob_nodes
        self.ci_level = ci_level

    def __repr__(self):
        e = self.edge_coef
        lines = ["", "Edge Coefficient Estimates (%d%% credible intervals)" % self.ci_level,
                 " node1 node2  estimate  lower_bound  upper_bound"]
        n = len(e["node1"])
        show = list(range(n)) if n <= 20 else list(range(10)) + [None] + list(range(n - 10, n))
        for i in show:
            if i is None:
                lines.append("   ...")
            else:
                lines.append(" %5d %5d  %8.3f  %11.3f  %11.3f" % (e["node1"][i], e["node2"][i], e["estimate"][i],
                                                                e["lower_bound"][i], e["upper_bound"][i]))
        lines.append("Node Probabilities")
        lines.extend(" %4d  %.3f" % (i + 1, p) for i, p in enumerate(self.prob_nodes["probability"]))
        return "\n".join(lines)


# -- index maps (src/utils.jl:17-57) -------------------------------------------------------------------
def lower_triangle(matrix):
    m = np.asarray(matrix)
    if m.ndim != 2 or m.shape[0] != m.shape[1]:
        raise ValueError("matrix must be square")
    V = m.shape[0]
    return np.concatenate([m[k:, k] for k in range(V)])


def create_lower_tri(vector, V):
    v = np.asarray(vector)
    mat = np.zeros((V, V), dtype=v.dtype)
    i = 0
    for k in range(V):
        mat[k:, k] = v[i:i + V - k]
        i += V - k
    return mat


def setup_X(X, x_transform=True):
    """setup_X! (src/gibbs.jl:239-247) -> n x q Float64 matrix."""
    if x_transform:
        return np.stack([lower_triangle(np.asarray(x, dtype=np.float64)) for x in X])
    return np.asarray(X, dtype=np.float64)


def _jround(x):
    """Julia round(): half to even."""
    return int(np.rint(x))


def _citation():
    return ("If you use BayesianNetworkRegression.jl, please cite:\n@article{Ozminkowski2022,\n"
            "author = {Ozminkowski, S. and Sol\\'{i}s-Lemus, C.},\nyear = {2022},\n"
            "title = {{Identifying microbial drivers in biological phenotypes with a Bayesian Network Regression model}},\n"
            "journal = {In preparation}\n}")


# -- run! : the row bookkeeping of src/gibbs.jl:849-864 turned into device run segments -----------------
def _run_rows(eng, first_index, nburn, total, purge_burn):
    """Generate rows exactly as run! would (1-based first_index/total, ring on purge_burn during burn-in)."""
    j = first_index
    seg_start, seg_len = j, 0
    for i in range(first_index, total + 1):
        seg_len += 1
        if purge_burn is not None and i < nburn and j == purge_burn + 1:
            eng.trace_row = seg_start - 1
            eng.run(seg_len)
            eng.copy_trace_rows(0, j - 1, 1)      # copy_table!(state, 1, j)
            j = 1
            seg_start, seg_len = 2, 0
        j += 1
    if seg_len:
        eng.trace_row = seg_start - 1
        eng.run(seg_len)


def _dist_world():
    """(rank, world) of an initialised torch.distributed process group, else (0, 1).  One process per GPU: every rank
    calls Fit with its own share of the chains (the reference spreads chains over Distributed.jl workers,
    src/gibbs.jl:946-948)."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return 0, 1


def _psrf(eng, nb, nsamp, streamed=False):
    """return_psrf_VOI (src/gibbs.jl:771-789): R-hat over table rows nb+1 .. nb+nsamp of every chain.
    streamed: the split-half moments of exactly those draws were accumulated while the chains ran
    (bnr_set_moment_window), so no chain but the first needs a trace.
    In a torch.distributed job the per-rank moments ([chain][half][param][mean, M2], 32 (V+q) bytes per chain -- the
    only data that crosses NVLink) are all-gathered and every rank reduces them in the same order, so all ranks take
    the same PSRF decisions."""
    if nsamp // 2 < 2:
        return np.full(eng.V, np.nan), np.full(eng.q, np.nan)
    if not streamed:
        eng.moments_from_trace(nb, nsamp)
    rank, world = _dist_world()
    if world == 1:
        return eng.rhat()
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", eng.device)
    _, cnt = eng.moments_device()
    mine = torch.empty(cnt, dtype=torch.float64, device=dev)
    eng.export_moments(mine.data_ptr())
    if dist.get_backend() == "nccl":
        allm = torch.empty(cnt * world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allm, mine)
    else:                                   # gloo (tests, CPU-side rendezvous): gather on the host
        host = torch.empty(cnt * world, dtype=torch.float64)
        dist.all_gather_into_tensor(host, mine.cpu())
        allm = host.to(dev)
    torch.cuda.synchronize(dev)
    return eng.rhat_from_moments(allm.data_ptr(), eng.C * world, eng.moment_half_len())


def _stream_last(eng, new_sweeps, nsamp):
    """Arm the streaming moments for the last nsamp of the next `new_sweeps` sweeps (the rows the PSRF will use)."""
    s_end = eng.iteration + new_sweeps
    eng.set_moment_window(s_end - nsamp + 1, nsamp)


def _fetch_state(eng, rows, what):
    st = Table()
    if what == "none":
        return st
    names = list(_UNI) if what == "full" else ["xi", "gamma"]
    for k in names:
        st[k] = eng.get_trace(0, k, 0, rows)
    return st


def _summary_ranks(nsamp, interval):
    """lw / hi of src/gibbs.jl:1221-1223 (Julia round = ties to even, 1-based indices)."""
    lower = (100 - interval) / 200.0
    return _jround(nsamp * lower), _jround(nsamp * (1.0 - lower))


def _device_summary(eng, nb, nsamp, interval=95):
    """Summary statistics of chain 1 computed on the GPU from the device traces (bnr_summary): no nsamp x q
    device-to-host copy and no host sort.  None when nsamp is too small for the interval (the reference raises
    a BoundsError inside Summary, not inside Fit!)."""
    lw, hi = _summary_ranks(nsamp, interval)
    if lw < 1 or hi > nsamp:
        return None
    gm, gl, gh, xm = eng.summary(0, nb, nsamp, lw, hi)
    return dict(interval=interval, mean=gm, lower=gl, upper=gh, xi_mean=xm, V=eng.V, q=eng.q)


def _max(a):
    return np.max(a) if len(a) else -np.inf   # NaN propagates like Julia's max(...)


def Fit(X, y, R, *, η=None, V=30, ζ=None, ι=None, aΔ=None, bΔ=None, ν=None, nburn=30000, nsamples=20000,
        mingen=0, maxgen=0, psrf_cutoff=1.01, x_transform=True, suppress_timer=False, num_chains=2, seed=None,
        purge_burn=None, filename="parameters.log", eta=None, zeta=None, iota=None, a_delta=None, b_delta=None,
        nu=None, device=0, return_state="full", verbose=False, chain_offset=0):
    """Drop-in for `Fit!(X, y, R; ...)` (src/gibbs.jl:725-751).  Greek keyword names are accepted as in the
    reference; ASCII aliases (eta, zeta, iota, a_delta, b_delta, nu) are equivalent.  Extra, engine-only
    keywords: device, return_state ("full" | "gamma_xi" | "none": how much of chain 1's table is copied back),
    chain_offset (global id of this GPU's first chain when several processes each fit a share of the chains;
    default rank * num_chains inside a torch.distributed job, where num_chains is the count PER RANK and the R-hat
    tables cover the chains of all ranks)."""
    def pick(greek, ascii_, default):
        return default if (greek is None and ascii_ is None) else (greek if greek is not None else ascii_)

    eta = pick(η, eta, 1.01)
    zeta = pick(ζ, zeta, 1.0)
    iota = pick(ι, iota, 1.0)
    a_delta = pick(aΔ, a_delta, 1.0)
    b_delta = pick(bΔ, b_delta, 1.0)
    nu = pick(ν, nu, 10)
    if seed is None:
        seed = random.randint(1, 55555)
    rank, world = _dist_world()
    if world > 1:
        if chain_offset == 0:
            chain_offset = rank * num_chains
        if rank != 0:
            filename = None                  # one parameters.log per job
    if filename:
        with open(filename, "w") as fh:
            fh.write("BayesianNetworkRegression.jl Fit! function\n")
            fh.write(datetime.datetime.now().strftime("%Y-%m-%d %H:%M:%S.%f")[:-3] + "\n")
            fh.write(_citation())
            fh.write("\n\nParameters:\n")
            fh.write("R=%s, η=%s, ζ=%s, ι=%s, aΔ=%s, bΔ=%s, ν=%s, nburn=%s, nsamples=%s, \n" %
                     (R, eta, zeta, iota, a_delta, b_delta, nu, nburn, nsamples))
            fh.write("mingen=%s, maxgen=%s, psrf_cutoff=%s, \n" % (mingen, maxgen, psrf_cutoff))
            fh.write("x_transform=%s, suppress_timer=%s, num_chains=%s, purge_burn=%s \n" %
                     (str(x_transform).lower(), str(suppress_timer).lower(), num_chains, purge_burn))
            fh.write("seed=%s" % seed)
    kw = dict(eta=eta, zeta=zeta, iota=iota, a_delta=a_delta, b_delta=b_delta, nu=nu,
              psrf_cutoff=psrf_cutoff, x_transform=x_transform, num_chains=num_chains, seed=seed,
              purge_burn=purge_burn, device=device, return_state=return_state, verbose=verbose,
              chain_offset=chain_offset)
    if mingen > 0 and maxgen > 0:
        return generate_samples_dbl(X, y, R, mingen=mingen, maxgen=maxgen, **kw)
    return generate_samples(X, y, R, nburn=nburn, nsamp=nsamples, maxburn=nburn + nsamples, **kw)


def _prepare(X, y, R, nu, x_transform):
    if nu < R:
        # the reference builds this ArgumentError without throwing it (src/gibbs.jl:901-902); the
        # InverseWishart draw is undefined for nu <= R-1, so the engine refuses instead of continuing.
        raise ValueError("ν value (%s) must be greater than R value (%s)" % (nu, R))
    if nu == R:
        print("Warning: ν==R may give poor accuracy. Consider increasing ν")
    Xn = setup_X(X, x_transform)
    y = np.asarray(y, dtype=np.float64).ravel()
    return Xn, y


def _normalise_purge(purge_burn, nburn):
    """src/gibbs.jl:930-936."""
    if purge_burn is not None and purge_burn < nburn and purge_burn != 0:
        if nburn % purge_burn != 0:
            purge_burn = purge_burn - (nburn % purge_burn)
        return purge_burn
    return None


def generate_samples(X, y, R, *, eta=1.01, zeta=1.0, iota=1.0, a_delta=1.0, b_delta=1.0, nu=10, nburn=30000,
                     nsamp=20000, maxburn=50000, psrf_cutoff=1.2, x_transform=True, num_chains=2, seed=None,
                     purge_burn=None, device=0, return_state="full", verbose=False, engine_hook=None,
                     chain_offset=0):
    """The "traditional" scheme (generate_samples!, src/gibbs.jl:897-1020)."""
    Xn, y = _prepare(X, y, R, nu, x_transform)
    total = nburn + nsamp
    purge_burn = _normalise_purge(purge_burn, nburn)
    tot_save = total if purge_burn is None else nsamp + purge_burn
    seed = random.randint(1, 55555) if seed is None else seed
    # R-hat needs the retained draws of EVERY chain.  Whenever those draws are always newly generated ones
    # (nburn >= nsamp, and the purge ring leaves room) their split-half moments are streamed on the device and only
    # chain 1 keeps a trace: memory is one chain's table instead of num_chains tables.
    streamed = nburn >= nsamp and nsamp >= 1 and (purge_burn is None or nsamp + purge_burn <= nburn)
    eng = Engine(Xn, y, R, num_chains=num_chains, seed=seed, device=device, chain_offset=chain_offset,
                 trace_rows=tot_save,
                 trace_full_chains=1 if return_state == "full" else 0, trace_gamma_xi_all=not streamed,
                 trace_gamma_xi_chains=1, eta=eta, zeta=zeta, iota=iota, a_delta=a_delta, b_delta=b_delta, nu=nu)
    try:
        eng.init_state()
        if streamed:
            _stream_last(eng, total - 1, nsamp)
        _run_rows(eng, 2, nburn, total, purge_burn)
        nb = purge_burn if purge_burn is not None else nburn
        tot_generated = nburn + nsamp
        rx, rg = _psrf(eng, nb, nsamp, streamed)
        if verbose:
            print("%d samples generated. Max PSRF XI: %.2f. Max PSRF Gamma: %.2f" % (tot_generated, _max(rx), _max(rg)))
        while (_max(rx) > psrf_cutoff or _max(rg) > psrf_cutoff) and tot_generated < maxburn + nsamp:
            if purge_burn is not None:
                num2move = 1 if nsamp + purge_burn <= nburn else nsamp + purge_burn - nburn
            else:
                num2move = total - nburn
            eng.copy_trace_rows(0, tot_save - num2move, num2move)
            a_total = num2move + nburn if num2move > 1 else nburn
            if streamed:
                _stream_last(eng, a_total - num2move, nsamp)
            _run_rows(eng, num2move + 1, (nburn - nsamp + num2move) if nburn > nsamp else 0, a_total, purge_burn)
            tot_generated += a_total - num2move
            rx, rg = _psrf(eng, nb, nsamp, streamed)
            if verbose:
                print("%d samples generated. Max PSRF XI: %.3f. Max PSRF Gamma: %.3f" %
                      (tot_generated, _max(rx), _max(rg)))
        if engine_hook is not None:
            engine_hook(eng)
        state = _fetch_state(eng, tot_save, return_state)
        extra = dict(status=eng.status(), tot_generated=tot_generated, seed=seed, gamma_mode=eng.gamma_mode,
                     device_summary=_device_summary(eng, nb, nsamp), rhat_streamed=streamed)
        return Results(state, rx, rg, nb, nsamp, extra)
    finally:
        eng.close()


def generate_samples_dbl(X, y, R, *, eta=1.01, zeta=1.0, iota=1.0, a_delta=1.0, b_delta=1.0, nu=10, mingen=10000,
                         maxgen=100000, psrf_cutoff=1.01, x_transform=True, num_chains=2, seed=None,
                         purge_burn=None, device=0, return_state="full", verbose=False, chain_offset=0):
    """The "doubling generation" scheme (generate_samples_dbl!, src/gibbs.jl:1051-1198).

    The retained window grows by mingen/2 draws per round, so its R-hat cannot come from one fixed streaming window.
    When mingen is a multiple of 4 every window and every split half is a whole number of blocks of mingen/4 sweeps:
    the device keeps per-block moments (bnr_set_moment_blocks) and merges them (bnr_moments_from_blocks), and only
    chain 1 keeps a trace.  Otherwise all chains are traced and R-hat is taken from the trace rows."""
    Xn, y = _prepare(X, y, R, nu, x_transform)
    nburn = _jround(mingen / 2)
    nsamp = mingen - nburn
    total = nburn + nsamp
    purge_burn = _normalise_purge(purge_burn, nburn)
    tot_save = total if purge_burn is None else nsamp + purge_burn
    halfburn = _jround(mingen / 2)
    rounds = max(0, math.ceil((maxgen - total) / max(mingen, 1)))
    capacity = max(tot_save, nsamp + (rounds + 1) * halfburn + halfburn)
    seed = random.randint(1, 55555) if seed is None else seed
    blocked = mingen % 4 == 0 and mingen >= 8 and purge_burn is None
    blk = mingen // 4
    eng = Engine(Xn, y, R, num_chains=num_chains, seed=seed, device=device, chain_offset=chain_offset,
                 trace_rows=capacity,
                 trace_full_chains=1 if return_state == "full" else 0, trace_gamma_xi_all=not blocked,
                 trace_gamma_xi_chains=1, eta=eta, zeta=zeta, iota=iota, a_delta=a_delta, b_delta=b_delta, nu=nu)
    try:
        if blocked:
            # block b = sweeps [b blk, (b+1) blk); after round k (k = 0: the first pass) mingen (k+1) sweeps exist and
            # the last mingen (k+1) / 2 of them are retained: blocks [2(k+1), 4(k+1))
            eng.set_moment_blocks(0, blk, 4 * (rounds + 1))
        eng.init_state()
        _run_rows(eng, 2, nburn, total, purge_burn)
        nb = purge_burn if purge_burn is not None else nburn
        tot_generated = total
        tot_samples = nsamp
        tot_sze = tot_save
        k = 0

        def psrf():
            if blocked:
                if (2 * (k + 1) * blk) // 2 < 2:
                    return np.full(eng.V, np.nan), np.full(eng.q, np.nan)
                eng.moments_from_blocks(2 * (k + 1), 2 * (k + 1))
                return _psrf(eng, nb, nsamp, streamed=True)
            return _psrf(eng, nb, nsamp)

        rx, rg = psrf()

        def unconverged():
            mx, mg = _max(rx), _max(rg)
            return mx > psrf_cutoff or mg > psrf_cutoff or np.isnan(mx) or np.isnan(mg)

        while unconverged() and tot_generated < maxgen:
            num2move = tot_samples
            tot_samples += halfburn
            nsamp = tot_samples
            new_save = tot_samples + halfburn
            eng.copy_trace_rows(0, tot_sze - num2move, num2move)
            _run_rows(eng, num2move + 1, 0, new_save, purge_burn)
            tot_sze = new_save
            tot_generated += mingen
            k += 1
            rx, rg = psrf()
            if verbose:
                print("%d samples generated. Max PSRF XI: %.3f. Max PSRF Gamma: %.3f" %
                      (tot_generated, _max(rx), _max(rg)))
        state = _fetch_state(eng, tot_sze, return_state)
        extra = dict(status=eng.status(), tot_generated=tot_generated, seed=seed, gamma_mode=eng.gamma_mode,
                     device_summary=_device_summary(eng, nb, nsamp), rhat_streamed=blocked)
        return Results(state, rx, rg, nb, nsamp, extra)
    finally:
        eng.close()


def Summary(results, interval=95, digits=3):
    """src/gibbs.jl:1214-1250: per-edge posterior mean and order-statistic credible bounds, per-node mean xi."""
    nburn, nsamp = results.burn_in, results.sampled
    total = nburn + nsamp
    lw, hi = _summary_ranks(nsamp, interval)
    if lw < 1 or hi > nsamp:
        raise IndexError("BoundsError: nsamp=%d too small for a %d%% interval" % (nsamp, interval))
    dev = (getattr(results, "extra", None) or {}).get("device_summary")
    if dev is not None and dev["interval"] == interval:
        # statistics were reduced on the GPU at the end of Fit (bnr_summary): only rounding and the tables remain
        mean, lo, up, xm, V = dev["mean"], dev["lower"], dev["upper"], dev["xi_mean"], dev["V"]
    else:
        # a different interval than the one reduced on the device: order statistics of chain 1's returned table
        if "gamma" not in results.state:
            raise ValueError("Summary(interval=%s) needs the gamma/xi table: call Fit with return_state='gamma_xi' "
                             "or 'full' (the device-side summary was computed for interval=%s only)"
                             % (interval, dev["interval"] if dev else None))
        g = np.asarray(results.state["gamma"])[nburn:total, :, 0]
        x = np.asarray(results.state["xi"])[nburn:total, :, 0]
        gs = np.sort(g, axis=0)
        mean, lo, up, xm = g.mean(axis=0), gs[lw - 1], gs[hi - 1], x.mean(axis=0)
        V = int((-1 + math.sqrt(1 + 8 * g.shape[1])) / 2)
    node1 = np.concatenate([np.full(V - k, k + 1, dtype=np.int64) for k in range(V)])
    node2 = np.concatenate([np.arange(k + 1, V + 1, dtype=np.int64) for k in range(V)])
    edge = dict(node1=node1, node2=node2, estimate=np.round(mean, digits),
                lower_bound=np.round(lo, digits), upper_bound=np.round(up, digits))
    nodes = dict(probability=np.round(xm, digits))
    return BNRSummary(edge, nodes, interval)
