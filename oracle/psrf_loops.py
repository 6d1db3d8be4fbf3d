"""CPU restatement of the reference's chain-orchestration loops  --  TEST INFRASTRUCTURE ONLY.

Only tests/ may import this module (the product path never does).  It restates, statement by statement and with the
reference's 1-based row indices, the host control flow around the Gibbs sweep:

  run!                   /root/reference/src/gibbs.jl:849-864    (row index j, purge_burn ring)
  initialize_and_run!    src/gibbs.jl:822-846                    (table size, first_index = 2)
  generate_samples!      src/gibbs.jl:897-1020                   ("extend burn-in" PSRF loop)
  generate_samples_dbl!  src/gibbs.jl:1051-1198                  ("doubling" PSRF loop)
  return_psrf_VOI        src/gibbs.jl:771-789                    (rows nburn+1 : nburn+nsamp of every chain)

The sweep itself is abstract: a *draw source* `draw(c, s)` returns whatever sweep number s of chain c produced (s = 0 is
the prior initialisation); because the engine keys its random numbers by (chain, sweep number), the draws of a long
reference run ARE the draws any bookkeeping scheme would see, so the loops can be checked exactly: which sweeps end up in
which table rows, how many sweeps are generated, which rows the R-hat sees and what it evaluates to.

Parity status: the reference's tests never exercise either PSRF loop with more than one pass (SURVEY 4.4 "not pinned");
this restatement is pinned to the source text only ("parity unpinned" at the golden-vector level).
"""
import numpy as np

from . import bnr_oracle as O


def julia_round(x):
    """Julia round(): half to even."""
    return int(np.rint(x))


class Tables:
    """states::Vector{Table}: per chain a list of rows; a row holds the SWEEP NUMBER whose state it stores (None =
    undef memory).  gibbs_sample!(state, j) reads row j-1 and writes row j (src/gibbs.jl:663-677)."""

    def __init__(self, num_chains):
        self.rows = [None] * num_chains
        self.sweeps = [0] * num_chains          # sweeps generated so far per chain (the RNG position)
        self.stale_reads = 0                    # gibbs_sample! calls whose row j-1 did NOT hold the chain's latest state

    def allocate(self, c, tot_save):
        self.rows[c] = [None] * tot_save

    def initialize_variables(self, c):
        self.rows[c][0] = 0                     # row 1 <- prior draw (src/gibbs.jl:191-224)
        self.sweeps[c] = 0

    def gibbs_sample(self, c, j):
        prev = self.rows[c][j - 2]
        if prev is None or prev != self.sweeps[c]:
            # the reference would continue the chain from an older (or uninitialised) row here; an engine that keeps
            # the chain state on the device always continues from the latest state.  Counted, so tests can tell.
            self.stale_reads += 1
        self.sweeps[c] += 1
        self.rows[c][j - 1] = self.sweeps[c]

    def copy_row(self, c, dst, src):            # copy_table!(state, dst, src)  (src/utils.jl:72-113)
        self.rows[c][dst - 1] = self.rows[c][src - 1]


def run(tab, c, first_index, nburn, total, purge_burn):
    """run! (src/gibbs.jl:849-864)."""
    j = first_index
    for i in range(first_index, total + 1):
        tab.gibbs_sample(c, j)
        if purge_burn is not None and i < nburn and j == purge_burn + 1:
            tab.copy_row(c, 1, j)
            j = 1
        j = j + 1


def normalise_purge(purge_burn, nburn):
    """src/gibbs.jl:930-936."""
    if purge_burn is not None and purge_burn < nburn and purge_burn != 0:
        if nburn % purge_burn != 0:
            purge_burn = purge_burn - (nburn % purge_burn)
        return purge_burn
    return None


def psrf_rows(tab, num_chains, nburn, nsamp):
    """The sweep numbers return_psrf_VOI hands to rhat: rows nburn+1 : nburn+nsamp of every chain (771-789)."""
    return [list(tab.rows[c][nburn:nburn + nsamp]) for c in range(num_chains)]


def rhat_of_rows(draw, rows):
    """rhat over (nsamp, nparams, num_chains) built from the draws of the given sweeps (src/convergence.jl:4-65).
    draw(c, s) -> (xi[V], gamma[q])."""
    xs = np.stack([np.stack([draw(c, s)[0] for s in r]) for c, r in enumerate(rows)], axis=2)
    gs = np.stack([np.stack([draw(c, s)[1] for s in r]) for c, r in enumerate(rows)], axis=2)
    return O.rhat(xs), O.rhat(gs)


def _max(v):
    """Julia max(v...): NaN propagates."""
    v = np.asarray(v, dtype=float)
    return np.nan if np.isnan(v).any() else (np.max(v) if v.size else -np.inf)


def generate_samples(draw, num_chains, nburn, nsamp, maxburn, psrf_cutoff, purge_burn=None, rhat_fn=None):
    """generate_samples! (src/gibbs.jl:897-1020).  Returns dict(rows, tot_generated, burn_in, sampled, rhat_xi,
    rhat_gamma, psrf_row_sweeps): `rows` = final table of every chain as sweep numbers.  rhat_fn(used) -> (rhat_xi,
    rhat_gamma) replaces the R-hat of the draws (used = per chain the sweep numbers of the rows the PSRF reads)."""
    if rhat_fn is None:
        rhat_fn = lambda used: rhat_of_rows(draw, used)
    total = nburn + nsamp
    purge_burn = normalise_purge(purge_burn, nburn)
    tab = Tables(num_chains)
    for c in range(num_chains):                                     # initialize_and_run! (822-846)
        tot_save = total if purge_burn is None else nsamp + purge_burn
        tab.allocate(c, tot_save)
        tab.initialize_variables(c)
        run(tab, c, 2, total - nsamp, total, purge_burn)
    tot_generated = nburn + nsamp
    nb = purge_burn if purge_burn is not None else nburn
    used = psrf_rows(tab, num_chains, nb, nsamp)
    rx, rg = rhat_fn(used)
    history = [(tot_generated, rx, rg, used)]
    # NaN > cutoff is false: a NaN R-hat ends the traditional loop
    while (_max(rx) > psrf_cutoff or _max(rg) > psrf_cutoff) and tot_generated < (maxburn + nsamp):
        if purge_burn is not None:
            num2move = 1 if nsamp + purge_burn <= nburn else nsamp + purge_burn - nburn
        else:
            num2move = total - nburn
        tot_sze = len(tab.rows[0])
        for c in range(num_chains):
            for i in range(1, num2move + 1):
                tab.copy_row(c, i, tot_sze - num2move + i)
            run(tab, c, num2move + 1, (nburn - nsamp + num2move) if nburn > nsamp else 0,
                (num2move + nburn) if num2move > 1 else nburn, purge_burn)
        A = (num2move + nburn) if num2move > 1 else nburn
        B = num2move
        tot_generated = tot_generated + A - B
        used = psrf_rows(tab, num_chains, nb, nsamp)
        rx, rg = rhat_fn(used)
        history.append((tot_generated, rx, rg, used))
    return dict(rows=tab.rows, tot_generated=tot_generated, burn_in=nb, sampled=nsamp, rhat_xi=rx, rhat_gamma=rg,
                psrf_row_sweeps=used, history=history, sweeps=list(tab.sweeps), stale_reads=tab.stale_reads)


def generate_samples_dbl(draw, num_chains, mingen, maxgen, psrf_cutoff, purge_burn=None, rhat_fn=None):
    """generate_samples_dbl! (src/gibbs.jl:1051-1198)."""
    if rhat_fn is None:
        rhat_fn = lambda used: rhat_of_rows(draw, used)
    nburn = julia_round(mingen / 2)
    nsamp = mingen - nburn
    total = nburn + nsamp
    purge_burn = normalise_purge(purge_burn, nburn)
    tab = Tables(num_chains)
    for c in range(num_chains):
        tot_save = total if purge_burn is None else nsamp + purge_burn
        tab.allocate(c, tot_save)
        tab.initialize_variables(c)
        run(tab, c, 2, total - nsamp, total, purge_burn)
    tot_generated = nburn + nsamp
    tot_samples = nsamp
    nb = purge_burn if purge_burn is not None else nburn
    used = psrf_rows(tab, num_chains, nb, nsamp)
    rx, rg = rhat_fn(used)
    history = [(tot_generated, rx, rg, used)]
    while (_max(rx) > psrf_cutoff or _max(rg) > psrf_cutoff or np.isnan(_max(rx)) or np.isnan(_max(rg))) \
            and tot_generated < maxgen:
        halfburn = julia_round(mingen / 2)
        num2move = tot_samples
        tot_samples = tot_samples + halfburn
        nsamp = tot_samples
        tot_sze = len(tab.rows[0])
        tot_save = tot_samples + halfburn
        for c in range(num_chains):
            old = tab.rows[c]
            tab.allocate(c, tot_save)                                # a fresh Table (1164-1170)
            for k in range(num2move):                                # copy_table!(state, states[c], 1:num2move, ...)
                tab.rows[c][k] = old[tot_sze - num2move + k]
            run(tab, c, num2move + 1, 0, tot_save, purge_burn)
        tot_generated = tot_generated + mingen
        used = psrf_rows(tab, num_chains, nb, nsamp)
        rx, rg = rhat_fn(used)
        history.append((tot_generated, rx, rg, used))
    return dict(rows=tab.rows, tot_generated=tot_generated, burn_in=nb, sampled=nsamp, rhat_xi=rx, rhat_gamma=rg,
                psrf_row_sweeps=used, history=history, sweeps=list(tab.sweeps), stale_reads=tab.stale_reads)
