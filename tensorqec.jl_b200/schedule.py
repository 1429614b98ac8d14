"""Lowering of a decoding factor graph to the static kernel schedule executed by `libtqec_cuda.so`.

What is lowered.  The reference builds, per decoder, a tensor network over binary labels
(src/decoding/tndecoder.jl:33-50 TNMAP; :97-146 TNMMAP/CSS; :186-238 TNMMAP/DEM) made of
  * prior factors over error variables (`single_qubit_tensor`, general_decoding.jl:5; `[1-p, p]`, tndecoder.jl:202-205),
  * one parity tensor per check / detector / logical row (`parity_check_matrix`, tndecoder.jl:42-44), whose last
    label is clamped by the syndrome (evidence, tndecoder.jl:54 / rank-1 vectors, :148-158) or left open
    (logical sectors, :134),
and hands it to OMEinsum for a pairwise contraction tree.  A parity tensor is the constraint
"xor of its labels = 0", i.e. it factorises exactly into a chain of rank-3 XOR tensors with 1-bit partial-parity
bonds -- the same factorisation the reference itself applies with `factorize=true` (`push_check_node!`,
tndecoder.jl:221-238).  Contracting that factorised network along a linear (caterpillar) tree that absorbs one
prior factor at a time gives the FRONTIER RECURRENCE below; its intermediate tensor is indexed by the partial
parities of the currently open checks (<= d+1 bits for a distance-d rotated surface code, versus 2(d-1) labels
for any tree over the unfactorised network).

Recurrence (semiring (+)/(x) = max/+ in the log domain for TNMAP, +/* for TNMMAP), for step t absorbing factor f
with variables v_1..v_r and table T[a], a = sum_j a_j 2^j (first label fastest, column-major like the reference):

    S_t[sigma'] = (+)_a  S_{t-1}[ sigma ]  (x)  T[a]      where sigma is sigma' with
        - every check c touched by the factor XOR-ed with the parity of the a_j of its variables,
        - checks whose first variable is in f ("opened") entering with partial parity 0,
        - checks whose last variable is in f ("closed") required to end at their syndrome bit and dropped,
        - observable rows (logical sectors) never closed: they index the final tensor.

Candidates are enumerated in ascending a; on exact FP64 ties the smallest a wins (documented tie rule).

Index layout used by the kernels.  The state is a dense array over w bits ("slots").  A step sees the FULL index:
slots 0..w_in-1 = the live checks in the order the previous step left them, slots w_in.. = the checks it opens.  Its
output index is any bit permutation of the surviving full slots (`perm[b]` = full slot of output bit b; closed slots
are not in the image).  For an output index the kernel (1) scatters its bits to their full slots and ORs in the shot's
syndrome values at the closed slots -> `full`, (2) reads the opened part `pat = full >> w_in`, which pins a coset
a0[pat] + ker of admissible candidates, (3) gathers S_in[(full ^ M[a]) & inmask] for a in that coset.  All tables are
emitted here; the permutation is chosen for the kernels (see the layout comment in `lower`).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

MAXPLUS = 0
SUMPROD = 1

HDR_INTS = 16
(H_R, H_WIN, H_NOPEN, H_NCLOSE, H_WOUT, H_NK, H_KB, H_OFF_T, H_OFF_ML, H_OFF_MK, H_OFF_A0, H_OFF_KER, H_OFF_VARS,
 H_OFF_CLOSE, H_RSV0, H_RSV1) = range(HDR_INTS)

MAX_FACTOR_RANK = 10
MAX_SMEM_WIDTH = 13            # 2 * 2^13 * 8 B = 128 KiB of ping-pong state per team
MAX_SUMPROD_ONCHIP_WIDTH = 11  # sum-product plans wider than this run on the global-memory executor (measured faster from 12 bits)
MAX_WIDE_WIDTH = 31            # the global-memory executor (wide.py, k_wide_pass): state of 2^w FP64 entries in HBM


@dataclass
class Factor:
    vars: Tuple[int, ...]
    table: np.ndarray            # flat, length 2^r, index a = sum_j a_j << j  (first variable fastest)


@dataclass
class Check:
    vars: Tuple[int, ...]
    kind: str                    # "syn" (clamped by syndrome bit `index`) | "obs" (kept open, output axis `index`)
    index: int


@dataclass
class Step:
    factor: int
    vars: Tuple[int, ...]
    w_in: int
    w_out: int
    opened: List[int]
    closed: List[Tuple[int, int]]            # (slot in the full index, syndrome bit), ascending slot
    perm: List[int]                          # perm[b] = slot in the full index of output bit b
    M: np.ndarray                            # (2^r,) masks in the full index space
    a0: np.ndarray                           # (2^n_open,) representative candidate per opened pattern, -1 = infeasible
    ker: np.ndarray                          # (nk,) kernel candidates, ascending
    table: np.ndarray                        # (2^r,) values in the semiring's domain (log for max-plus)
    quad: bool = False                       # two factors absorbed at once in the 16-output block form


@dataclass
class Schedule:
    semiring: int
    n_vars: int
    n_checks: int
    n_obs: int
    steps: List[Step]
    obs_slot: List[int]
    order: List[int]
    factors: List[Factor]
    checks: List[Check]
    w_max: int = 0
    cost: float = 0.0                        # sum_t 2^w_out * nk  = candidate evaluations per shot
    log2_scale: int = 0                      # sum-product: the emitted tables are the factors times 2^-e_f; sum of e_f
    # flat encoding for the C-ABI
    hdr: Optional[np.ndarray] = None
    ints: Optional[np.ndarray] = None
    tables: Optional[np.ndarray] = None

    def ops_per_shot(self):
        """Algorithmic FP64 operations per shot: one (x) per candidate and one (+) per candidate beyond the first."""
        mul = sum((1 << s.w_out) * len(s.ker) for s in self.steps)
        add = sum((1 << s.w_out) * (len(s.ker) - 1) for s in self.steps)
        return mul, add


# ------------------------------------------------------------------------------------------------------------
def libm_log(table) -> np.ndarray:
    """Log-weights of a factor through the C library's `log`, one entry at a time (log 0 = -inf).  numpy's vectorised
    log (SIMD kernels chosen by CPU features) differs from it in the last bit on ~0.3 % of arguments; the C++ lowering
    inside libtqec_cuda.so (csrc/tqec_lower.cpp) calls the same libm, so both lowerings emit identical tables."""
    return np.array([math.log(x) if x > 0.0 else -math.inf for x in np.asarray(table, dtype=np.float64).reshape(-1)],
                    dtype=np.float64)


def flat_table(t) -> np.ndarray:
    """Tensor with one axis per variable (reference layout, column-major) -> flat array indexed by a."""
    t = np.asarray(t, dtype=np.float64)
    return t.reshape(-1, order="F").copy()


def merge_overlapping(factors: Sequence[Factor], n_vars: int, check_vars, allow_negative: bool = False) -> List[Factor]:
    """Make the prior factors a partition of the variables: multiply factors that share a variable into one
    (exact; the reference contracts the same product), add an all-ones factor for variables no prior mentions
    (TensorInference's per-variable unity tensor)."""
    parent = list(range(len(factors)))

    def find(i):
        while parent[i] != i:
            parent[i] = parent[parent[i]]
            i = parent[i]
        return i

    owner = {}
    for i, f in enumerate(factors):
        if len(set(f.vars)) != len(f.vars):
            raise ValueError(f"factor {i} repeats a variable: {f.vars}")
        for v in f.vars:
            if v in owner:
                parent[find(i)] = find(owner[v])
            else:
                owner[v] = i
    groups = {}
    for i in range(len(factors)):
        groups.setdefault(find(i), []).append(i)
    out = []
    for root in sorted(groups, key=lambda r: min(groups[r])):
        members = groups[root]
        if len(members) == 1:
            f = factors[members[0]]
            out.append(Factor(tuple(f.vars), np.asarray(f.table, dtype=np.float64).copy()))
            continue
        vs = []
        for i in members:
            for v in factors[i].vars:
                if v not in vs:
                    vs.append(v)
        if len(vs) > MAX_FACTOR_RANK:
            raise ValueError(f"overlapping prior factors merge into rank {len(vs)} > {MAX_FACTOR_RANK}")
        tab = np.ones(1 << len(vs))
        a = np.arange(1 << len(vs))
        for i in members:
            f = factors[i]
            idx = np.zeros_like(a)
            for j, v in enumerate(f.vars):
                idx |= ((a >> vs.index(v)) & 1) << j
            tab = tab * np.asarray(f.table)[idx]
        out.append(Factor(tuple(vs), tab))
    covered = {v for f in out for v in f.vars}
    for v in sorted(set(check_vars) - covered):
        out.append(Factor((v,), np.ones(2)))
    for f in out:
        if len(f.vars) > MAX_FACTOR_RANK:
            raise ValueError(f"prior factor of rank {len(f.vars)} > {MAX_FACTOR_RANK} is not supported")
        if f.table.shape != (1 << len(f.vars),):
            raise ValueError("factor table must have 2^rank entries")
        if not np.isfinite(f.table).all() or ((f.table < 0).any() and not allow_negative):
            raise ValueError("prior factor entries must be finite and non-negative")
    return out


def map_order(original: Sequence[Factor], merged: Sequence[Factor], order: Sequence[int]) -> List[int]:
    """A caller's absorption order refers to ITS prior tensors (e.g. the leaf order of an OMEinsum tree); the schedule
    absorbs the merged factors (overlapping priors multiplied, unity factors appended).  Each merged factor takes the
    place of the first of its members in the caller's order; factors the caller does not know come last."""
    order = [int(i) for i in order]
    if sorted(order) != list(range(len(original))):
        raise ValueError("order must be a permutation of the prior tensors")
    home = {}
    for mi, f in enumerate(merged):
        for v in f.vars:
            home[v] = mi
    out: List[int] = []
    for i in order:
        for v in original[i].vars:
            mi = home[v]
            if mi not in out:
                out.append(mi)
    out += [mi for mi in range(len(merged)) if mi not in out]
    return out


# ------------------------------------------------------------------------------------------------------------
class _Sim:
    """Incremental frontier simulation used by the ordering heuristic."""

    def __init__(self, factors, checks):
        self.factors = factors
        self.checks = checks
        self.f_checks = [[] for _ in factors]                     # checks touched by each factor
        var_owner = {v: i for i, f in enumerate(factors) for v in f.vars}
        self.c_factors = []
        for ci, c in enumerate(checks):
            fs = sorted({var_owner[v] for v in c.vars})
            self.c_factors.append(fs)
            for fi in fs:
                self.f_checks[fi].append(ci)
        self.var_owner = var_owner

    def step_cost(self, fi, remaining, live):
        """(w_out, log2 work) of absorbing factor fi given `remaining[c]` = #factors of check c not yet absorbed."""
        f = self.factors[fi]
        opened = [c for c in self.f_checks[fi] if c not in live]
        n_close = sum(1 for c in self.f_checks[fi] if remaining[c] == 1 and self.checks[c].kind == "syn")
        w_out = len(live) + len(opened) - n_close
        # rank of the opened part
        rows = []
        for v in f.vars:
            m = 0
            for k, c in enumerate(opened):
                if v in self.checks[c].vars:
                    m |= 1 << k
            rows.append(m)
        rank = _gf2_rank(rows)
        return w_out, w_out + len(f.vars) - rank


def _gf2_rank(rows):
    rows = [r for r in rows if r]
    rank = 0
    while rows:
        p = max(rows)
        hb = p.bit_length() - 1
        rows = [r ^ p if (r >> hb) & 1 else r for r in rows if r != p]
        rows = [r for r in rows if r]
        rank += 1
    return rank


def _evaluate(order, sim):
    remaining = [len(fs) for fs in sim.c_factors]
    live = set()
    wmax, cost = 0, 0.0
    for fi in order:
        w_out, lw = sim.step_cost(fi, remaining, live)
        for c in sim.f_checks[fi]:
            live.add(c)
        for c in sim.f_checks[fi]:
            remaining[c] -= 1
            if remaining[c] == 0 and sim.checks[c].kind == "syn":
                live.discard(c)
        wmax = max(wmax, w_out, w_out)
        cost += 2.0 ** lw
    return wmax, cost


def _score(wmax, cost, n_steps):
    """Candidate evaluations per shot plus the per-step set-up a team pays once per pass (~290 candidate-equivalents,
    measured on B200), shared by the 2^(10 - w_max) shots a team packs into one pass."""
    return cost + 290.0 * n_steps / (1 << max(0, 10 - wmax))


def _greedy_order(sim, start):
    nF = len(sim.factors)
    remaining = [len(fs) for fs in sim.c_factors]
    live = set()
    done = [False] * nF
    order = []
    cand = {start}
    while len(order) < nF:
        best = None
        pool = cand if cand else {i for i in range(nF) if not done[i]}
        for fi in pool:
            w_out, lw = sim.step_cost(fi, remaining, live)
            key = (w_out, lw, fi)
            if best is None or key < best[0]:
                best = (key, fi)
        fi = best[1]
        order.append(fi)
        done[fi] = True
        cand.discard(fi)
        for c in sim.f_checks[fi]:
            live.add(c)
        for c in sim.f_checks[fi]:
            remaining[c] -= 1
            if remaining[c] == 0 and sim.checks[c].kind == "syn":
                live.discard(c)
        # neighbours: factors sharing a live check
        for c in sim.f_checks[fi]:
            for fj in sim.c_factors[c]:
                if not done[fj]:
                    cand.add(fj)
    return order


def spectral_orders(sim) -> List[List[int]]:
    """Sweep orders from the Fiedler vector of the check graph (two checks are adjacent when a factor touches both):
    the second eigenvector of its Laplacian is a smooth coordinate along the longest extent of the graph, and absorbing
    the factors in the order of the largest coordinate among their checks sweeps a front across it.  For the 3-D
    detector graphs of circuit-level models this finds fronts far narrower than the local greedy search (d = 5 x 5
    rounds surface-code memory: 29 bits against 35).  Observable rows are left out (they touch factors all along a
    logical operator and would short-circuit the graph); connected components are swept one after the other."""
    nF = len(sim.factors)
    ids = [c for c, ch in enumerate(sim.checks) if ch.kind == "syn"]
    if len(ids) < 3:
        return []
    loc = {c: k for k, c in enumerate(ids)}
    n = len(ids)
    A = np.zeros((n, n))
    for fc in sim.f_checks:
        fc = [loc[c] for c in fc if c in loc]
        for a in fc:
            for b in fc:
                if a != b:
                    A[a, b] = 1.0
    comp = [-1] * n
    comps = []
    for s0 in range(n):
        if comp[s0] >= 0:
            continue
        comp[s0] = len(comps)
        stack, members = [s0], []
        while stack:
            u = stack.pop()
            members.append(u)
            for v in np.flatnonzero(A[u]):
                if comp[v] < 0:
                    comp[v] = len(comps)
                    stack.append(int(v))
        comps.append(sorted(members))
    x = np.zeros(n)
    base = 0.0
    for members in comps:
        if len(members) >= 3:
            sub = A[np.ix_(members, members)]
            _, vecs = np.linalg.eigh(np.diag(sub.sum(axis=1)) - sub)
            v = vecs[:, 1]
            if v[int(np.argmax(np.abs(v)))] < 0:                  # eigenvectors are defined up to a sign: fix it
                v = -v
            v = v - v.min()
        else:
            v = np.arange(len(members), dtype=np.float64)
        x[members] = base + v
        base += float(v.max()) + 1.0
    out = []
    for xx in (x, -x):
        hi = [max((xx[loc[c]] for c in sim.f_checks[i] if c in loc), default=0.0) for i in range(nF)]
        out.append(sorted(range(nF), key=lambda i: (hi[i], i)))
    return out


def choose_order(factors, checks, max_starts=24):
    """Pick the absorption order: the best (by candidate evaluations per shot) of the natural order, its reverse,
    greedy minimum-frontier sweeps from several starting factors, and the spectral sweeps."""
    sim = _Sim(factors, checks)
    nF = len(factors)
    cands = [list(range(nF)), list(range(nF - 1, -1, -1))]
    if nF > 600:
        max_starts = 6                                            # the greedy search is quadratic in the factor count
    deg = sorted(range(nF), key=lambda i: (len(sim.f_checks[i]), i))
    starts = list(dict.fromkeys(deg[: max_starts // 2] + [0, nF - 1] + list(range(0, nF, max(1, nF // (max_starts // 2))))))
    for s in starts[:max_starts]:
        cands.append(_greedy_order(sim, s))
    scored = [(_evaluate(o, sim), i) for i, o in enumerate(cands)]
    (wmax, cost), i = min(scored, key=lambda x: (_score(x[0][0], x[0][1], nF), x[0][0], x[1]))
    if wmax > MAX_SMEM_WIDTH:
        # too wide for the on-chip kernels: the plan will run on the global-memory executor, where only the width and
        # the candidate count matter; add the spectral sweeps (never consulted for plans that fit on chip, so their
        # orders -- and the kernels' tie-breaks -- are unchanged)
        n0 = len(cands)
        cands += spectral_orders(sim)
        scored += [(_evaluate(o, sim), n0 + k) for k, o in enumerate(cands[n0:])]
        (wmax, cost), i = min(scored, key=lambda x: (x[0][0], x[0][1], x[1]))
    return cands[i]


# ------------------------------------------------------------------------------------------------------------
def lower(factors: Sequence[Factor], checks: Sequence[Check], semiring: int, n_vars: int, n_checks: int,
          n_obs: int = 0, order: Optional[Sequence[int]] = None, max_width: int = MAX_SMEM_WIDTH,
          fuse: Optional[bool] = None, _split=None, stable: bool = False) -> Schedule:
    """Factor graph -> `Schedule` (see module docstring).  `order` optionally fixes the absorption order of the
    (merged) factors, e.g. the leaf order of a contraction tree chosen by the caller's optimiser.  `stable` keeps the
    surviving checks in their relative order and puts opened checks on top (a monotone `perm`): the layout of the
    global-memory executor (wide.py), where a bit permutation is data movement instead of address arithmetic."""
    all_check_vars = {v for c in checks for v in c.vars}
    original = list(factors)
    # signed factors (Clifford-network inference: Pauli-representation tensors have entries +-1) are legal for sum-product
    factors = merge_overlapping(original, n_vars, all_check_vars, allow_negative=semiring == SUMPROD)
    checks = [Check(tuple(dict.fromkeys(c.vars)), c.kind, c.index) for c in checks]
    for c in checks:
        if c.kind not in ("syn", "obs"):
            raise ValueError(f"unknown check kind {c.kind!r}")
    if order is None:
        order = choose_order(factors, checks)
    order = list(order)
    if len(order) == len(original) and len(original) != len(factors):
        order = map_order(original, factors, order)              # the caller's order refers to its own prior tensors
    if sorted(order) != list(range(len(factors))):
        raise ValueError("order must be a permutation of the (merged) factors")
    if fuse is None:
        import os as _os
        fuse = semiring == MAXPLUS and _os.environ.get("TQEC_NO_FUSE") is None
    if stable:
        fuse = False
    if fuse and _split is None:
        # absorb consecutive factor pairs as one step where that yields the 16-output block form: greedy pairing along
        # the order; a pair that does not come out in the canonical form is forbidden and the pairing redone from there
        forbidden = set()
        best = None
        for _ in range(4 * len(order) + 4):
            mf, morder, pairs = _merge_pairs(factors, order, forbidden)
            if not any(p is not None for p in pairs):
                break
            trial = lower(mf, checks, semiring, n_vars, n_checks, n_obs, order=morder, max_width=max_width, fuse=True,
                          _split=True)
            bad = [pairs[i] for i, st in enumerate(trial.steps) if pairs[i] is not None and not (st.quad and st.w_out == 9)]
            if not bad:
                best = trial
                break
            forbidden.add(bad[0])
        if best is not None and any(st.quad for st in best.steps):
            return best
    sim = _Sim(factors, checks)
    remaining = [len(fs) for fs in sim.c_factors]

    # pass 1: which checks every step touches / opens / closes (depends on the order only)
    orphan = [ci for ci, fs in enumerate(sim.c_factors) if not fs]
    plan = []
    seen = set()
    for t, fi in enumerate(order):
        touched = list(sim.f_checks[fi]) + (orphan if t == 0 else [])
        opened = [c for c in touched if c not in seen]
        seen.update(opened)
        closing = []
        for c in touched:
            if c in orphan:
                if checks[c].kind == "syn":
                    closing.append(c)
                continue
            remaining[c] -= 1
            if remaining[c] == 0 and checks[c].kind == "syn":
                closing.append(c)
        plan.append((touched, opened, closing))

    # pass 2: slot layout.  The output bit order of a step is free (the gather realises any bit permutation at no run
    # time cost), so it is chosen for the kernels: the five low bits (the lane bits of a warp team) hold checks the
    # next steps leave alone, the checks flipped by this step's kernel candidates sit right above them (both outputs of
    # such a pair then belong to one thread), the checks the NEXT step closes sit above those (a closed bit among the
    # four low bits of the input index would cost 2-way shared-memory bank conflicts), opened checks take the top.
    live: List[int] = []
    steps: List[Step] = []
    cost = 0.0
    wmax = 0
    log2_scale = 0
    log2_run = 0.0
    for t, fi in enumerate(order):
        f = factors[fi]
        r = len(f.vars)
        touched, opened, closing = plan[t]
        w_in = len(live)
        full = live + opened
        if len(full) > MAX_WIDE_WIDTH:
            raise ValueError(f"frontier needs {len(full)} bits > {MAX_WIDE_WIDTH}: no executor holds such a state "
                             f"(choose a sweep-like absorption order)")
        pos = {c: k for k, c in enumerate(full)}
        closed = sorted((pos[c], checks[c].index) for c in closing)
        m = []
        for v in f.vars:
            mv = 0
            for c in sim.f_checks[fi]:
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
            a0[pat[a]] = a                                   # smallest a of each coset
        ker = np.flatnonzero(pat == 0).astype(np.int64)      # ascending, ker[0] = 0
        kept_old = [c for c in live if c not in closing]
        kmask = 0
        for k in ker:
            kmask |= int(M[k])
        km = [c for c in kept_old if (kmask >> pos[c]) & 1]
        # "quad" steps (two factors absorbed at once: 4 candidates per output, 2 checks opened): if the candidates that
        # leave every closed check alone form a 4-element group that realises all 4 opened patterns and each of its two
        # generators flips exactly one surviving check, those two checks go to output bits 5, 6 and serve as coset
        # representatives: a block of 16 outputs (4 values of those bits x 4 opened patterns) then reads exactly 16
        # inputs, each used by 4 outputs (see quad_step in csrc/tqec_decode.cu).
        quad = False
        q5 = q6 = None
        if len(ker) == 4 and n_open == 2 and r <= 6:
            cmask = 0
            for slot, _ in closed:
                cmask |= 1 << slot
            inmask = (1 << w_in) - 1
            kc = [a for a in range(1 << r) if (int(M[a]) & cmask) == 0]
            reps = {int(pat[a]): a for a in kc}
            if len(kc) == 4 and len(reps) == 4:
                c1, c2 = int(M[reps[1]]) & inmask, int(M[reps[2]]) & inmask
                ok1 = c1 == 0 or (c1 & (c1 - 1)) == 0
                ok2 = c2 == 0 or (c2 & (c2 - 1)) == 0
                if ok1 and ok2 and (c1 or c2) and c1 != c2:
                    q5 = full[c1.bit_length() - 1] if c1 else None      # a generator may flip no surviving check:
                    q6 = full[c2.bit_length() - 1] if c2 else None      # that output bit then holds any other check
                    if all(q is None or q in kept_old for q in (q5, q6)):
                        quad = True
                        for pp in range(4):
                            a0[pp] = reps[pp]
                        km = [q for q in (q5, q6) if q is not None]
        nxt = set(plan[t + 1][2]) if t + 1 < len(order) else set()
        cn = [c for c in kept_old if c in nxt and c not in km]
        others = [c for c in kept_old if c not in km and c not in cn]
        if stable:
            quad = False
            out_old = kept_old
        elif quad and len(others) + len(cn) >= 5 + (q5 is None) + (q6 is None):
            pool = others + cn                                   # bits 0-4, fillers for an unused generator bit, rest
            low, pool = pool[:5], pool[5:]
            b5 = q5 if q5 is not None else pool.pop(0)
            b6 = q6 if q6 is not None else pool.pop(0)
            rest_cn = [c for c in pool if c in cn]
            out_old = low + [b5, b6] + rest_cn + [c for c in pool if c not in cn]
        elif len(others) >= 5:
            quad = False if quad else quad
            out_old = others[:5] + km + cn + others[5:]
        else:
            quad = False
            out_old = others + cn + km
        live = out_old + [c for c in opened if c not in closing]
        perm = [pos[c] for c in live]
        w_out = len(live)
        if semiring == MAXPLUS:
            tab = libm_log(f.table)
        else:
            # static per-step scaling: every factor is multiplied by a power of two (exact in FP64) chosen so that the
            # running product of the factors' largest entries stays within [2^-1/2, 2^1/2] -- the state can never
            # overflow however many factors there are (1605 at d = 5 x 5 rounds circuit level) and starts from the best
            # place against underflow; the exponents are summed and re-applied by the host after the decode
            tab = f.table.copy()
            mx = float(np.abs(tab).max())
            if mx > 0.0:
                log2_run += math.log2(mx)
                e = int(np.rint(log2_run))
                log2_run -= e
                tab = np.ldexp(tab, -e)
                log2_scale += e
        steps.append(Step(fi, tuple(f.vars), w_in, w_out, opened, closed, perm, M, a0, ker, tab, quad))
        cost += float(1 << w_out) * len(ker)
        wmax = max(wmax, w_in, w_out)                        # the full index is never materialised
    if wmax > max_width:
        raise ValueError(f"frontier needs {wmax} bits > {max_width}: the schedule does not fit the on-chip state")
    obs_slot = [-1] * n_obs
    for k, c in enumerate(live):
        if checks[c].kind != "obs":
            raise AssertionError("a clamped check survived the sweep")
        obs_slot[checks[c].index] = k
    if any(s < 0 for s in obs_slot) or len(live) != n_obs:
        raise ValueError("every observable row must be declared exactly once")
    sch = Schedule(semiring, n_vars, n_checks, n_obs, steps, obs_slot, order, factors, list(checks), wmax, cost, log2_scale)
    _encode(sch)
    return sch


def _merge_pairs(factors, order, forbidden=()):
    """Greedy pairing along the order: two consecutive factors are absorbed as one (product table over the union of
    their variables) unless that pair is forbidden.  -> (factors, order, pairs) with pairs[i] = the two original factor
    ids merged into new factor i, or None."""
    out, pairs = [], []
    k = 0
    while k < len(order):
        a = order[k]
        b = order[k + 1] if k + 1 < len(order) else None
        if b is not None and (a, b) not in forbidden and len(factors[a].vars) + len(factors[b].vars) <= 4:
            f, g = factors[a], factors[b]
            r = len(f.vars)
            idx = np.arange(1 << (r + len(g.vars)))
            out.append(Factor(tuple(f.vars) + tuple(g.vars), f.table[idx & ((1 << r) - 1)] * g.table[idx >> r]))
            pairs.append((a, b))
            k += 2
        else:
            out.append(factors[a])
            pairs.append(None)
            k += 1
    return out, list(range(len(out))), pairs


def _encode(s: Schedule):
    """Flatten to the arrays `tqec_plan_create` ingests (include/tqec.h, `tqec_plan_desc`)."""
    hdr = np.zeros((len(s.steps), HDR_INTS), dtype=np.int32)
    ints: List[int] = []
    tabs: List[float] = []
    zero = -np.inf if s.semiring == MAXPLUS else 0.0
    for t, st in enumerate(s.steps):
        r = len(st.vars)
        n_open = len(st.opened)
        nk = len(st.ker)
        inmask = (1 << st.w_in) - 1
        h = hdr[t]
        h[H_R], h[H_WIN], h[H_NOPEN], h[H_NCLOSE], h[H_WOUT] = r, st.w_in, n_open, len(st.closed), st.w_out
        h[H_NK], h[H_KB] = nk, nk.bit_length() - 1
        assert nk & (nk - 1) == 0
        h[H_OFF_T] = len(tabs)
        for p in range(1 << n_open):
            for k in range(nk):
                tabs.append(float(st.table[st.a0[p] ^ st.ker[k]]) if st.a0[p] >= 0 else zero)
        h[H_OFF_ML] = len(ints)
        ints += [int(st.M[st.a0[p]] & inmask) if st.a0[p] >= 0 else 0 for p in range(1 << n_open)]
        h[H_OFF_MK] = len(ints)
        ints += [int(st.M[k] & inmask) for k in st.ker]
        h[H_OFF_A0] = len(ints)
        ints += [int(st.a0[p]) if st.a0[p] >= 0 else 0 for p in range(1 << n_open)]
        h[H_OFF_KER] = len(ints)
        ints += [int(k) for k in st.ker]
        h[H_OFF_VARS] = len(ints)
        ints += [int(v) for v in st.vars]
        h[H_OFF_CLOSE] = len(ints)
        for slot, bit in st.closed:
            ints += [int(slot), int(bit)]
        ints += [int(x) for x in st.perm]                    # perm[w_out] follows the closed list
    s.hdr = hdr
    s.ints = np.asarray(ints if ints else [0], dtype=np.int32)
    s.tables = np.asarray(tabs if tabs else [0.0], dtype=np.float64)
