"""Lowering for the global-memory executor (`k_wide_pass`, csrc/tqec_wide.cu): frontier plans whose state does not fit
on chip (14 <= w_max <= 31 bits; 2^w FP64 entries per shot live in HBM).

Why it exists.  The detector graph of a circuit-level memory experiment is three-dimensional and depolarizing noise
couples its X and Z halves, so any exact contraction of the reference's DEM network (src/decoding/tndecoder.jl:186-238)
carries a 2-D cross-section: 29 bits at d = 5 x 5 rounds for the frontier recurrence (BASELINE configs[3]); the
reference's own tree has the same width.  The recurrence is the one of schedule.py; what changes is where the state lives
and how a step reaches it.

  * STATE IN HBM, ping-pong: two arrays of `n_batch * 2^w_cap` doubles; entry sigma of shot b at (b << w_cap) | sigma.
  * STABLE LAYOUT.  Surviving checks keep their relative order, opened checks go on top, closed checks drop out and
    the bits above them move down.  The oldest live checks -- the ones about to be closed, i.e. the ones the next steps
    touch -- therefore sit in the lowest index bits.
  * PASSES.  Consecutive steps are grouped into a pass.  A pass touches a set of checks (the TILE bits: every check one
    of its steps flips, opens or closes, plus the `low_bits` lowest index bits before and after the pass so that
    global loads and stores come in contiguous runs of 2^low_bits entries); all other live checks are SPECTATORS.  For
    every value of the spectator bits (and every shot of the batch) one CTA loads the 2^t_in tile entries into shared
    memory, runs the pass's steps there (ping-pong, the generic gather of schedule.py in tile-local coordinates) and
    stores the 2^t_out entries of the result: ONE read and ONE write of the state per pass instead of per step.  The
    kernel is HBM-bound by construction: 16 * 2^w bytes of traffic per pass.
  * Everything the kernel needs per step is tabulated here in tile-local coordinates (same fields as the step header
    of include/tqec.h; the output permutation is monotone, so it is given as the mask of surviving full slots).

Sum-product plans only (TNMMAP, DEM or CSS); max-plus plans of this width would additionally stream back-pointers.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import schedule as S

T_MAX = 12               # tile bits: 2 * 2^12 * 8 B = 64 KiB of ping-pong state per CTA, three CTAs per SM
LOW_BITS = 4             # contiguous run of global loads / stores: 2^4 entries = 128 B
MAX_PASS_STEPS = 240
PASS_INTS = 16           # pass header (int32)
STEP_INTS = 16           # local step header (int32)
(P_WIN, P_WOUT, P_TIN, P_TOUT, P_NSTEPS, P_STEP0, P_TINMASK, P_TOUTMASK, P_OFF_INTS, P_N_INTS, P_OFF_TAB, P_N_TAB) = range(12)
(L_WIN, L_NOPEN, L_NCLOSE, L_WOUT, L_NK, L_OFF_T, L_OFF_ML, L_OFF_MK, L_OFF_CLOSE, L_KEEPMASK) = range(10)


@dataclass
class LocalStep:
    factor: int
    w_in: int
    n_open: int
    w_out: int
    nk: int
    closed: List[Tuple[int, int]]              # (slot in the local full index, syndrome bit)
    keepmask: int                              # local full slots that survive (output bit b = b-th set bit)
    ML: np.ndarray                             # (2^n_open,) in-state mask of the coset representative
    MK: np.ndarray                             # (nk,) in-state mask of kernel candidate k
    T: np.ndarray                              # (2^n_open, nk) factor values


@dataclass
class WidePass:
    t0: int
    t1: int
    w_in: int
    w_out: int
    tin_mask: int                              # positions of the tile bits in the global index before the pass
    tout_mask: int                             # ... after the pass
    t_in: int
    t_out: int
    t_peak: int
    steps: List[LocalStep]


@dataclass
class WidePlan:
    semiring: int
    n_vars: int
    n_checks: int
    n_obs: int
    w_cap: int                                 # widest global state between passes (bits): what lives in HBM
    w_max: int                                 # widest state at any step, inside passes too (frontier width of the order)
    t_max: int
    passes: List[WidePass]
    obs_pos: List[int]                         # position of observable i in the final index
    log2_scale: int
    order: List[int]
    factors: List[S.Factor] = field(default_factory=list)      # merged factors / checks the plan was lowered from
    checks: List[S.Check] = field(default_factory=list)
    cost: float = 0.0                          # candidate evaluations per shot
    bytes_per_shot: float = 0.0                # HBM traffic per shot: 8 * sum_pass (2^w_in + 2^w_out)
    # flat tables for the C ABI
    pass_hdr: Optional[np.ndarray] = None
    step_hdr: Optional[np.ndarray] = None
    ints: Optional[np.ndarray] = None
    tables: Optional[np.ndarray] = None


def _roles(factors, checks, order):
    owner = {v: i for i, f in enumerate(factors) for v in f.vars}
    c_factors = [sorted({owner[v] for v in c.vars}) for c in checks]
    if any(not fs for fs in c_factors):
        raise ValueError("wide lowering: a check without variables (orphan) is not supported")
    f_checks = [[] for _ in factors]
    for ci, fs in enumerate(c_factors):
        for fi in fs:
            f_checks[fi].append(ci)
    remaining = [len(x) for x in c_factors]
    seen = set()
    out = []
    for fi in order:
        touched = f_checks[fi]
        opened = [c for c in touched if c not in seen]
        seen.update(opened)
        closing = []
        for c in touched:
            remaining[c] -= 1
            if remaining[c] == 0 and checks[c].kind == "syn":
                closing.append(c)
        out.append((fi, touched, opened, closing))
    return out


def _local_step(live, f, fi, touched, opened, closing, checks, table):
    """One step in the coordinates of `live` (ordered check list) -> (LocalStep, new live list).  Same algebra as
    schedule.lower pass 2 with the stable layout."""
    r = len(f.vars)
    w_in = len(live)
    full = live + opened
    pos = {c: k for k, c in enumerate(full)}
    m = []
    for v in f.vars:
        mv = 0
        for c in touched:
            if v in checks[c].vars:
                mv |= 1 << pos[c]
        m.append(mv)
    A = np.arange(1 << r)
    M = np.zeros(1 << r, dtype=np.int64)
    for j in range(r):
        M ^= np.where((A >> j) & 1, m[j], 0)
    n_open = len(opened)
    pat = M >> w_in
    a0 = np.full(1 << n_open, -1, dtype=np.int64)
    for a in range((1 << r) - 1, -1, -1):
        a0[pat[a]] = a
    ker = np.flatnonzero(pat == 0).astype(np.int64)
    nk = len(ker)
    inmask = (1 << w_in) - 1
    cl = set(closing)
    keep = [c for c in full if c not in cl]
    keepmask = 0
    for c in keep:
        keepmask |= 1 << pos[c]
    closed = sorted((pos[c], checks[c].index) for c in closing)
    ML = np.array([int(M[a0[p]]) & inmask if a0[p] >= 0 else 0 for p in range(1 << n_open)], dtype=np.int64)
    MK = np.array([int(M[k]) & inmask for k in ker], dtype=np.int64)
    T = np.zeros((1 << n_open, nk))
    for p in range(1 << n_open):
        if a0[p] >= 0:
            for k in range(nk):
                T[p, k] = table[a0[p] ^ ker[k]]
    return LocalStep(fi, w_in, n_open, len(keep), nk, closed, keepmask, ML, MK, T), keep


def lower_wide(factors: Sequence[S.Factor], checks: Sequence[S.Check], semiring: int, n_vars: int, n_checks: int,
               n_obs: int, order: Optional[Sequence[int]] = None, t_max: int = T_MAX, low_bits: int = LOW_BITS,
               max_drop_bits: float = 0.0) -> WidePlan:
    """Factor graph -> `WidePlan`.  Factors are merged / completed as in schedule.lower; `order` as there.
    `max_drop_bits` > 0 (dynamic rescaling, which acts between passes): a pass ends before the product of its steps'
    largest / smallest non-zero factor entries exceeds 2^max_drop_bits, so no shot can underflow inside one pass."""
    if semiring != S.SUMPROD:
        raise ValueError("the global-memory executor runs sum-product plans only")
    all_check_vars = {v for c in checks for v in c.vars}
    factors = S.merge_overlapping(list(factors), n_vars, all_check_vars, allow_negative=True)
    checks = [S.Check(tuple(dict.fromkeys(c.vars)), c.kind, c.index) for c in checks]
    if order is None:
        order = S.choose_order(factors, checks)
    order = list(order)
    if sorted(order) != list(range(len(factors))):
        raise ValueError("order must be a permutation of the (merged) factors")
    roles = _roles(factors, checks, order)
    n = len(roles)
    # static power-of-two scaling of every factor (schedule.lower does the same): exact, undone by the host
    tabs = []
    log2_scale = 0
    log2_run = 0.0
    for fi, *_ in roles:
        tab = np.asarray(factors[fi].table, dtype=np.float64).copy()
        mx = float(np.abs(tab).max())
        if mx > 0.0:
            log2_run += math.log2(mx)                             # running product of the maxima stays near 1
            e = int(np.rint(log2_run))
            log2_run -= e
            tab = np.ldexp(tab, -e)
            log2_scale += e
        tabs.append(tab)
    drops = []
    for tab in tabs:
        nz = np.abs(tab[tab != 0])
        drops.append(float(np.log2(nz.max() / nz.min())) if nz.size else 0.0)

    def simulate(t0, t1, glive, lb):
        """-> None if steps t0..t1-1 do not fit one tile, else (tile checks, global live after, local peak width)."""
        touched_all = set()
        g = list(glive)
        for t in range(t0, t1):
            _, touched, opened, closing = roles[t]
            touched_all.update(touched)
            cl = set(closing)
            g = [c for c in g + opened if c not in cl]
        tile = touched_all | set(glive[:lb]) | set(g[:lb])
        w = sum(1 for c in glive if c in tile)
        peak = w
        for t in range(t0, t1):
            _, touched, opened, closing = roles[t]
            peak = max(peak, w + len(opened))                     # opened checks coexist with the ones this step closes
            w = w + len(opened) - len(closing)
            if w + len(closing) > S.MAX_WIDE_WIDTH:
                return None
        if peak > t_max:
            return None
        if max_drop_bits > 0 and t1 - t0 > 1 and sum(drops[t0:t1]) > max_drop_bits:
            return None
        return tile, g, peak

    # plans made of rank-1 factors only (detector error models) keep a pass within what one butterfly block holds
    max_pass_steps = 96 if all(len(f.vars) == 1 for f in factors) else MAX_PASS_STEPS
    passes: List[WidePass] = []
    glive: List[int] = []
    w_cap = 0
    w_peak = 0
    cost = 0.0
    traffic = 0.0
    t = 0
    while t < n:
        best = None
        for lb in range(low_bits, -1, -1):
            for t1 in range(t + 1, min(n, t + max_pass_steps) + 1):
                r = simulate(t, t1, glive, lb)
                if r is None:
                    break
                best = (t1, r)
            if best is not None:
                break
        if best is None:
            raise ValueError(f"wide lowering: step {t} alone needs more than {t_max} tile bits")
        t1, (tile, gout, peak) = best
        # spare tile bits go to the lowest untouched spectators: larger tiles, longer contiguous runs in HBM
        room = t_max - peak
        for c in glive:
            if room <= 0:
                break
            if c not in tile:
                tile.add(c)
                room -= 1
        tin_mask = sum(1 << k for k, c in enumerate(glive) if c in tile)
        tout_mask = sum(1 << k for k, c in enumerate(gout) if c in tile)
        L = [c for c in glive if c in tile]
        t_in = len(L)
        lsteps = []
        for tt in range(t, t1):
            fi, touched, opened, closing = roles[tt]
            ls, L = _local_step(L, factors[fi], fi, touched, opened, closing, checks, tabs[tt])
            lsteps.append(ls)
        if len(glive) > S.MAX_WIDE_WIDTH or len(gout) > S.MAX_WIDE_WIDTH:
            raise ValueError(f"frontier needs {max(len(glive), len(gout))} bits > {S.MAX_WIDE_WIDTH}")
        assert L == [c for c in gout if c in tile]
        n_spec = len(glive) - t_in
        for ls in lsteps:
            cost += float(1 << (n_spec + ls.w_out)) * ls.nk
            w_peak = max(w_peak, n_spec + ls.w_out)
        traffic += 8.0 * ((1 << len(glive)) + (1 << len(gout)))
        passes.append(WidePass(t, t1, len(glive), len(gout), tin_mask, tout_mask, t_in, len(L), peak, lsteps))
        w_cap = max(w_cap, len(glive), len(gout))
        glive = gout
        t = t1
    obs_pos = [-1] * n_obs
    for k, c in enumerate(glive):
        if checks[c].kind != "obs":
            raise AssertionError("a clamped check survived the sweep")
        obs_pos[checks[c].index] = k
    if any(p < 0 for p in obs_pos) or len(glive) != n_obs:
        raise ValueError("every observable row must be declared exactly once")
    plan = WidePlan(semiring, n_vars, n_checks, n_obs, w_cap, w_peak, t_max, passes, obs_pos, log2_scale, order, factors, list(checks),
                    cost, traffic)
    _encode(plan)
    return plan


def _encode(p: WidePlan):
    n_steps = sum(len(ps.steps) for ps in p.passes)
    ph = np.zeros((len(p.passes), PASS_INTS), dtype=np.int32)
    sh = np.zeros((n_steps, STEP_INTS), dtype=np.int32)
    ints: List[int] = []
    tabs: List[float] = []
    s = 0
    for i, ps in enumerate(p.passes):
        h = ph[i]
        h[P_WIN], h[P_WOUT], h[P_TIN], h[P_TOUT], h[P_NSTEPS], h[P_STEP0] = ps.w_in, ps.w_out, ps.t_in, ps.t_out, len(ps.steps), s
        h[P_TINMASK], h[P_TOUTMASK] = ps.tin_mask, ps.tout_mask
        i0, f0 = len(ints), len(tabs)
        for ls in ps.steps:
            q = sh[s]
            q[L_WIN], q[L_NOPEN], q[L_NCLOSE], q[L_WOUT], q[L_NK] = ls.w_in, ls.n_open, len(ls.closed), ls.w_out, ls.nk
            # offsets are relative to the pass's block of the pools (the kernel copies that block to shared memory)
            q[L_OFF_T] = len(tabs) - f0
            tabs += [float(x) for x in ls.T.reshape(-1)]
            q[L_OFF_ML] = len(ints) - i0
            ints += [int(x) for x in ls.ML]
            q[L_OFF_MK] = len(ints) - i0
            ints += [int(x) for x in ls.MK]
            q[L_OFF_CLOSE] = len(ints) - i0
            for slot, bit in ls.closed:
                ints += [int(slot), int(bit)]
            q[L_KEEPMASK] = ls.keepmask                       # output index -> full index: deposit at these bits
            s += 1
        h[P_OFF_INTS], h[P_N_INTS], h[P_OFF_TAB], h[P_N_TAB] = i0, len(ints) - i0, f0, len(tabs) - f0
    p.pass_hdr, p.step_hdr = ph, sh
    p.ints = np.asarray(ints if ints else [0], dtype=np.int32)
    p.tables = np.asarray(tabs if tabs else [0.0], dtype=np.float64)
