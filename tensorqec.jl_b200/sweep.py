"""In-place patch sweep: second lowering of the frontier recurrence, executed by `k_sweep` (csrc/tqec_sweep.cu).

The recurrence is the one of schedule.py (one factor absorbed per step, candidates in ascending assignment order, strict
`>` so the smallest assignment wins exact ties).  What changes is how a team of 32 threads walks it:

  * STATIC SLOTS.  Every open check owns one bit position ("slot") of the state index for its whole life.  In the bulk
    of a planar sweep a step closes one check and opens another through the same variable; the opened check inherits
    the slot of the closed one, so the state never has to be permuted and is updated IN PLACE (no ping-pong copy: half
    the shared memory per team, twice the resident warps).
  * PATCHES.  A step only touches the slots of its factor's checks (<= 4).  A thread loads the 2^M state entries that
    differ in those M slots (a "patch") into registers, absorbs one or two factors on them with static register
    indices ("layers": in-place butterflies  Out[j] = max_k R[j ^ F(k)] + T[p(j)][k]), and stores them back.  The other
    index bits are spread over the 32 lanes and a short loop.  The closed checks' syndrome bits are folded into the load
    address, so register indices are relative to the shot's syndrome.
  * HEAD TABLE.  The first steps (narrow states, checks opened into fresh slots) depend on a handful of syndrome bits
    only; their result is tabulated at compile time for every value of those bits (state + partial configuration per
    entry) and a pass starts by copying one table row.
  * LATE FOLD.  A check opened by the first layer of a super-step and closed by the second cannot be folded into the
    load address (its slot still holds the first layer's closed check); its syndrome bit is applied to the second
    layer's table rows (a row-swapped copy of the table, chosen per shot) and to the store address instead.
  * FRESH PINS.  A variable may also open a check without closing one (the column turn of an even-distance code): the
    opened check then takes a DEAD slot (a chain whose check closed without a successor, or a new one) and the pinned
    variable's flip mask contains its own bit, so that output bit 1 reads the live (bit 0) half of the slot.
  * PINNED / FREE VARIABLES.  A variable that opens a check is pinned to that check's output bit (the opened check
    inherits the slot of a check the same variable closes) and may flip further patch bits when it is 1; a variable
    that opens nothing is a candidate dimension (butterfly over its flip mask).  Observable checks of sum-product
    plans are never closed: their slots stay live and index the output marginals.
  * The register-level wiring of a super-step (patch size, pinned bits, flip masks) must be one of the shapes compiled
    into the kernel (`MENU`, mirrored by csrc/tqec_sweep_menu.h).  Plans with any other step shape are not lowered
    (`lower_sweep` returns None) and run on the general kernels of csrc/tqec_decode.cu.

Reference semantics: the same contraction as schedule.py (tndecoder.jl:33-57, 97-165, 186-253), evaluated in the same
order with the same IEEE operations as the unfused schedule, so results are bit-identical to it.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

from . import schedule as S

NB = 10                 # index bits of one team pass: W slot bits + sg shot bits (1024 FP64 entries = 8 KiB)
MAX_PATCH = 4           # patch bits (16 registers of state)
MAX_HEAD_BITS = 6       # syndrome bits the head table may depend on
REC_INTS = 32           # forward record per super-step (int32)
TB_INTS = 64            # traceback record per super-step (int32)

# Register-level shapes compiled into k_sweep: (M, ((pinned bits), (free masks)[, (extra flip masks of the pinned
# variables)]) per layer).  Canonical form = the
# lexicographically smallest descriptor over all orderings of the patch bits.  Keep in sync with tqec_sweep_menu.h
# (tests/test_sweep_cpu.py compares the two).
MENU: List[tuple] = [
    (2, (((), (1, 2)),)),
    (2, (((0,), (2,)),)),
    (2, (((0,), (2,)), ((), (1, 2)))),
    (2, (((0,), (2,)), ((), (2, 1)))),
    (2, (((0,), (2,)), ((0,), (2,)))),
    (3, (((), (1, 2)), ((0,), (4,)))),
    (3, (((), (3, 4)),)),
    (3, (((0,), (6,)),)),
    (3, (((0,), (2,)), ((), (5, 2)))),
    (3, (((0,), (2,)), ((1,), (5,)))),
    (3, (((0,), (6,)), ((1,), (1,)))),
    (4, (((0,), (6,)), ((1,), (9,)))),
    # shapes with FRESH pins (a pinned variable whose flip mask contains its own bit: the opened check takes a dead slot)
    # and the shapes around them: even-distance and rectangular rotated surface codes
    (3, (((0, 1), (), (5, 2)),)),
    (3, (((), (1, 6)),)),
    (4, (((0,), (6,)), ((3,), (2,)))),
    (3, (((0, 1), (), (5, 2)), ((), (1, 2)))),
    (4, (((0,), (2,)), ((2,), (10,)))),
    (3, (((0, 1), (), (5, 2)), ((0,), (2,)))),
    (3, (((), (1, 2)), ((1,), (4,)))),
    (3, (((0,), (2,)), ((), (2, 5)))),
    # sum-product only
    (4, (((0,), (14,)),)),
    (4, (((0,), (6,), (8,)),)),
    (4, (((0,), (2,)), ((1,), (5,), (8,)))),
    (4, (((0,), (2,)), ((2,), (10,), (1,)))),
    (4, (((0,), (6,)), ((), (9, 6)))),
    # sum-product only, extended instantiation: TNMMAP plans of even-distance and rectangular codes
    (4, (((), (3, 12)),)),
    (3, (((0, 1), (), (1, 6)),)),
    (4, (((0, 1), (), (1, 6)), ((2,), (9,)))),
    (4, (((0, 1), (), (1, 6)), ((), (9, 6)))),
    (3, (((0,), (6,)), ((0,), (6,)))),
    (4, (((0,), (6,)), ((), (6, 9)))),
    (3, (((0,), (6,)), ((), (6, 1)))),
]


MENU_MAXPLUS = 20       # shapes 0..19 are compiled into the max-plus kernel; later ones are sum-product only
_DISCOVER = None


def phys(x: int) -> int:
    """Shared-memory swizzle of a 10-bit logical entry index (entries are 8 bytes): index bits 4..7 are XOR-ed into
    bits 0..3, so 32 lanes whose indices differ in any 4 positions with distinct residues mod 4 hit 16 distinct bank
    pairs per half warp.  Linear over GF(2): phys(a ^ b) = phys(a) ^ phys(b)."""
    return x ^ ((x >> 4) & 15)


@dataclass
class Layer:
    step: int                                  # index into the unfused schedule
    factor: int
    vars: Tuple[int, ...]
    pinned: List[Tuple[int, int]]              # (variable index j in the factor, patch bit)
    free: List[Tuple[int, int]]                # (variable index j, patch-local flip mask)
    closed: List[Tuple[int, int]]              # (syndrome bit, patch bit)
    T: np.ndarray = None                       # [2^NP * 2^NF] values: index pidx * 2^NF + k
    pk: List[int] = field(default_factory=list)  # per pinned variable: patch bits it flips besides its own (checks it
                                               # touches that stay open or close without being re-opened)


@dataclass
class SuperStep:
    layers: List[Layer]
    chains: List[int]                          # chain (slot) ids of patch bits 0..M-1
    menu: int = -1
    pos: List[int] = field(default_factory=list)        # logical position of patch bit b
    lanepos: List[int] = field(default_factory=list)    # 5 positions spanned by the lane id
    looppos: List[int] = field(default_factory=list)    # active positions walked by the iteration loop
    conflict: bool = False
    late: Optional[Tuple[int, int]] = None     # (syndrome bit, patch bit): check opened by layer 0 and closed by layer 1
    fresh: List[List[int]] = field(default_factory=list)      # per layer: chains that come alive (fresh pins)
    wbase: int = 0                             # first back-pointer word (per lane) of the step inside a pass
    bpp: int = 0                               # back-pointer bits per patch
    n_words: int = 0


@dataclass
class SweepPlan:
    semiring: int
    n_vars: int
    n_checks: int
    n_obs: int
    W: int
    sg: int
    head_steps: int
    head_bits: List[int]                       # syndrome bits the head depends on (bit j of the head pattern)
    head_state: np.ndarray                     # (2^nh, 2^W) FP64, index = slot bits
    head_cfg: np.ndarray                       # (2^nh, 2^W, ncw) uint64 partial configurations (max-plus)
    ssteps: List[SuperStep]
    bp_words: int                              # back-pointer words per lane per pass
    out_index: List[int]                       # logical slot index of final entry i (2^n_obs entries; [0] for max-plus)
    conflicts: int = 0
    # flat device tables
    rec: Optional[np.ndarray] = None           # (n_ss, REC_INTS) int32
    tb: Optional[np.ndarray] = None            # (n_ss, TB_INTS) int32
    lanetab: Optional[np.ndarray] = None       # (n_ss, 32) uint32: phys lane address | lane sub bits << 16
    tvals: Optional[np.ndarray] = None         # pooled FP64 layer tables

    def instr_estimate(self):
        return sum((1 << (NB - 5)) * (4 + 6 * len(s.layers)) + 40 for s in self.ssteps)


# ------------------------------------------------------------------------------------------------------------------
def _roles(factors, checks, order):
    owner = {v: i for i, f in enumerate(factors) for v in f.vars}
    c_factors = [sorted({owner[v] for v in c.vars}) for c in checks]
    if any(not fs for fs in c_factors):
        return None                                            # orphan checks: general kernels only
    remaining = [len(x) for x in c_factors]
    f_checks = [[] for _ in factors]
    for ci, fs in enumerate(c_factors):
        for fi in fs:
            f_checks[fi].append(ci)
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


def _classify(f, touched, opened, closing, checks):
    """-> (pinned [(j, opened check, closed check, other checks)], free [(j, touched checks)]) or None if the step cannot
    run in place.  A PINNED variable touches exactly one opened check (which no other variable touches) and a closing
    check whose slot the opened one inherits; the other checks it touches are flipped when it is 1.  A FREE variable
    touches no opened check."""
    opened, closing = set(opened), set(closing)
    if opened & closing:
        return None
    pinned, free = [], []
    donors = set()
    for j, v in enumerate(f.vars):
        tv = [c for c in touched if v in checks[c].vars]
        ov = [c for c in tv if c in opened]
        if not ov:
            free.append((j, tv))
            continue
        if len(ov) != 1 or sum(1 for w in f.vars if w in checks[ov[0]].vars) != 1:
            return None
        rest = [c for c in tv if c != ov[0]]
        cand = [c for c in rest if c in closing and c not in donors]
        if not cand:
            # FRESH pin: the opened check takes a dead slot (donor None); its own slot joins the flip mask of the pinned
            # variable, so that output bit 1 reads the live (bit 0) half of the slot
            pinned.append((j, ov[0], None, rest))
            continue
        donors.add(cand[0])
        pinned.append((j, ov[0], cand[0], [c for c in rest if c != cand[0]]))
    if len(pinned) > 2 or len(free) > 2:
        return None
    return pinned, free


def _descriptor(layers_raw, chain_order):
    """Descriptor of a group for a given ordering of its patch chains: per layer (pinned bits, free masks[, masks of the
    other bits each pinned variable flips -- only when some are non-zero])."""
    bit = {ch: b for b, ch in enumerate(chain_order)}
    desc = []
    for pinned, free, pk in layers_raw:
        pb = tuple(bit[ch] for _, ch in pinned)
        fm = tuple(sum(1 << bit[ch] for ch in chs) for _, chs in free)
        pm = tuple(sum(1 << bit[ch] for ch in chs) for chs in pk)
        desc.append((pb, fm, pm) if any(pm) else (pb, fm))
    return (len(chain_order), tuple(desc))


def _canonical(layers_raw, chains):
    best = None
    for perm in itertools.permutations(chains):
        d = _descriptor(layers_raw, perm)
        if best is None or d < best[0]:
            best = (d, list(perm))
    return best


def _head_eval(sch, h, roles, head_bits, live_order):
    """Tabulate the first h steps for every value of the syndrome bits they close.
    -> state (2^nh, 2^W) in slot order `live_order` (bit k of the index = parity of check live_order[k]) and, for
    max-plus, the partial configuration of every entry as packed words (uint64 (2^nh, 2^W, ceil(n_vars/64))).
    A closed check's axis is not indexed by a known syndrome bit but kept as a batch axis ("value of that syndrome
    bit"), so the batch grows only as bits are closed and the early steps are evaluated once."""
    maxplus = sch.semiring == S.MAXPLUS
    nh = len(head_bits)
    ncw = max(1, (sch.n_vars + 63) // 64)
    batch: List[int] = []                           # syndrome bits closed so far = leading axes, in closing order
    axes: List[int] = []                            # open checks = trailing axes
    St = np.full((), 0.0 if maxplus else 1.0)
    cfg = np.zeros((ncw,), dtype=np.uint64) if maxplus else None
    zero = -np.inf if maxplus else 0.0
    for t in range(h):
        fi, touched, opened, closing = roles[t]
        f = sch.factors[fi]
        T = sch.steps[t].table
        for c in opened:
            St = np.stack([St, np.full_like(St, zero)], axis=-1)
            if maxplus:
                cfg = np.stack([cfg, cfg], axis=-2)
            axes.append(c)
        nb = len(batch)
        best = None
        for a in range(1 << len(f.vars)):
            flips = []
            for c in touched:
                p = 0
                for j, v in enumerate(f.vars):
                    if v in sch.checks[c].vars:
                        p ^= (a >> j) & 1
                if p:
                    flips.append(c)
            ax = tuple(nb + axes.index(c) for c in flips)
            src = np.flip(St, axis=ax) if ax else St
            cand = src + T[a] if maxplus else src * T[a]
            if maxplus:
                # every variable belongs to exactly one factor, so its bit is still clear: OR in the assignment
                amask = np.zeros(ncw, dtype=np.uint64)
                for j, v in enumerate(f.vars):
                    if (a >> j) & 1:
                        amask[v >> 6] |= np.uint64(1) << np.uint64(v & 63)
                csrc = (np.flip(cfg, axis=ax) if ax else cfg) | amask
            if best is None:
                best = cand.copy()
                if maxplus:
                    bcfg = csrc
            elif maxplus:
                upd = cand > best                              # strict: the smallest assignment wins exact ties
                best = np.where(upd, cand, best)
                bcfg = np.where(upd[..., None], csrc, bcfg)
            else:
                best = best + cand
        St = best
        if maxplus:
            cfg = bcfg
        for c in closing:
            # the check's parity must equal its syndrome bit: its axis becomes the batch axis of that bit
            k = nb + axes.index(c)
            St = np.moveaxis(St, k, nb)
            if maxplus:
                cfg = np.moveaxis(cfg, k, nb)
            axes.remove(c)
            batch.append(sch.checks[c].index)
            nb += 1
    assert batch == list(head_bits) and sorted(axes) == sorted(live_order)
    W = len(live_order)
    # head pattern bit j = batch axis j, state index bit k = check live_order[k]: slowest axis first for a C reshape
    permu = list(range(nh - 1, -1, -1)) + [nh + axes.index(c) for c in reversed(live_order)]
    St = np.ascontiguousarray(np.transpose(St, permu)).reshape(1 << nh, 1 << W)
    if maxplus:
        cfg = np.ascontiguousarray(np.transpose(cfg, permu + [cfg.ndim - 1])).reshape(1 << nh, 1 << W, ncw)
    return St, cfg


def _assign_positions(groups, W, seed=0):
    """chain -> position (0..W-1) minimising the number of super-steps whose patch contains positions q and q+4
    (q < 4): those lose a bank-pair bit and pay 2-way shared-memory conflicts."""
    def cost(p):
        c = 0
        for chains in groups:
            ps = {p[ch] for ch in chains}
            if any(q in ps and q + 4 in ps for q in range(4)):
                c += 1
        return c

    state = (seed + 1) & 0xFFFFFFFFFFFFFFFF

    def shuffled():
        # Fisher-Yates driven by a 64-bit LCG (Knuth's MMIX constants): the same few lines in csrc/tqec_lower_sweep.cpp
        nonlocal state
        q = list(range(W))
        for i in range(W - 1, 0, -1):
            state = (state * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
            j = (state >> 33) % (i + 1)
            q[i], q[j] = q[j], q[i]
        return q

    best_p, best_c = list(range(W)), None
    for restart in range(40):
        p = shuffled() if restart else list(range(W))
        c = cost(p)
        improved = True
        while improved and c:
            improved = False
            for i in range(W):
                for j in range(i + 1, W):
                    p[i], p[j] = p[j], p[i]
                    c2 = cost(p)
                    if c2 < c:
                        c, improved = c2, True
                    else:
                        p[i], p[j] = p[j], p[i]
        if best_c is None or c < best_c:
            best_p, best_c = list(p), c
        if c == 0:
            break
    return [int(x) for x in best_p], best_c


# ------------------------------------------------------------------------------------------------------------------
def lower_sweep(sch: S.Schedule, max_head_bits: int = MAX_HEAD_BITS) -> Optional[SweepPlan]:
    """Unfused `Schedule` (one factor per step) -> `SweepPlan`, or None when the plan does not fit the in-place form."""
    if any(st.quad for st in sch.steps) or len(sch.steps) != len(sch.factors):
        raise ValueError("lower_sweep expects the unfused schedule (schedule.lower(..., fuse=False))")
    factors, checks, order = sch.factors, sch.checks, sch.order
    roles = _roles(factors, checks, order)
    if roles is None:
        return None
    n = len(roles)
    cls = [_classify(factors[fi], touched, opened, closing, checks) for fi, touched, opened, closing in roles]
    # head = the shortest prefix after which every step runs in place, extended while the head table stays small
    h = n
    while h > 0 and cls[h - 1] is not None:
        h -= 1
    if h == 0:
        h = 1
    head_bits = [checks[c].index for t in range(h) for c in roles[t][3]]
    if len(head_bits) > max_head_bits or h >= n:
        return None
    while h + 1 < n and len(head_bits) + len(roles[h][3]) <= max_head_bits:
        head_bits += [checks[c].index for c in roles[h][3]]
        h += 1
    # live checks after the head, in opening order -> chains 0..W-1
    live: List[int] = []
    for t in range(h):
        _, _, opened, closing = roles[t]
        live += opened
        live = [c for c in live if c not in closing]
    W = len(live)
    if W > NB or W < 1:
        return None
    chain_of = {c: k for k, c in enumerate(live)}
    live_order = list(live)
    W0 = W                                                     # chains alive after the head; fresh pins may add more

    # layers in chain coordinates
    raw = []
    dead_since: dict = {}                                      # chain -> step after which it is free again
    for t in range(h, n):
        fi, touched, opened, closing = roles[t]
        pinned, free = cls[t]
        f = factors[fi]
        fresh_chains = []
        for j, o, c, _ in pinned:
            if c is None:
                # a chain that died in an earlier step (such two steps never share a super-step: see match), else a new one
                cands = sorted(ch for ch, ts in dead_since.items() if ts + 1 <= t)
                if cands:
                    ch = cands[0]
                    del dead_since[ch]
                else:
                    ch = W
                    W += 1
                    if W > NB:
                        return None
                chain_of[o] = ch
                fresh_chains.append(ch)
        lp = [(j, chain_of[c] if c is not None else chain_of[o]) for j, o, c, _ in pinned]
        lk = [[chain_of[c] for c in extra] + ([chain_of[o]] if c0 is None else []) for _, o, c0, extra in pinned]
        lf = [(j, [chain_of[c] for c in tv]) for j, tv in free]
        lc = [(checks[c].index, chain_of[c]) for c in closing]
        chains = sorted({ch for _, ch in lp} | {ch for chs in lk for ch in chs} | {ch for _, chs in lf for ch in chs} |
                        {ch for _, ch in lc})
        donors_used = {chain_of[c] for _, _, c, _ in pinned if c is not None}
        for j, o, c, _ in pinned:
            if c is not None:
                chain_of[o] = chain_of[c]
        for c in closing:
            if chain_of[c] not in donors_used:
                dead_since[chain_of[c]] = t
        freed = [chain_of[c] for c in closing if chain_of[c] not in donors_used]
        raw.append(dict(step=t, fi=fi, pinned=lp, pk=lk, free=lf, closed=lc, chains=chains, fresh=fresh_chains, freed=freed))
        if len(chains) > MAX_PATCH or len(chains) == 0:
            return None
    sg = NB - W
    if sg > 5 or sg < 0:
        return None
    # an observable must still be alive at the end, everything else dead
    final_live = [c for c in chain_of if checks[c].kind == "obs"]
    menu_ix = {m: i for i, m in enumerate(MENU) if sch.semiring == S.SUMPROD or i < MENU_MAXPLUS}
    if _DISCOVER is not None:                                  # development aid: record the shapes a plan would need
        class _Any(dict):
            def __contains__(self, d):
                _DISCOVER.setdefault(d, 0)
                _DISCOVER[d] += 1
                return True

            def __getitem__(self, d):
                return dict.get(self, d, -1)
        menu_ix = _Any(menu_ix)

    allow_late = True

    def match(group):
        chains = sorted({ch for g in group for ch in g["chains"]})
        if len(chains) > MAX_PATCH:
            return None
        if len(group) == 2 and set(group[0]["freed"]) & set(group[1]["fresh"]):
            return None                                        # a slot cannot die and reopen inside one super-step
        if len(group) == 2:
            # a check opened by the first layer and closed by the second cannot be folded into the load address (its
            # slot still holds the first layer's closed check): its syndrome bit is applied late -- to the second
            # layer's table rows and to the store address (see SuperStep.late); at most one such check per super-step
            both = {ch for _, ch in group[0]["pinned"]} & {ch for _, ch in group[1]["closed"]}
            if len(both) > 1 or (both and not allow_late):
                return None
        if sch.semiring == S.MAXPLUS and sum(len(g["free"]) for g in group) << len(chains) > 32:
            return None                                        # back-pointers of a patch must fit one 32-bit word
        d, perm = _canonical([(g["pinned"], g["free"], g["pk"]) for g in group], chains)
        if d not in menu_ix:
            return None
        return menu_ix[d], perm

    # pairing: fewest super-steps (dynamic programme over the step list; a pair must match a menu shape)
    nr = len(raw)
    single = [match(raw[k:k + 1]) for k in range(nr)]
    pair = [match(raw[k:k + 2]) if k + 1 < nr else None for k in range(nr)]
    INF = 10 ** 9
    best = [0] * (nr + 2)
    best[nr + 1] = INF
    take = [0] * nr
    for k in range(nr - 1, -1, -1):
        best[k] = INF
        if single[k] is not None and 1 + best[k + 1] < best[k]:
            best[k], take[k] = 1 + best[k + 1], 1
        if pair[k] is not None and 1 + best[k + 2] <= best[k]:
            best[k], take[k] = 1 + best[k + 2], 2
    if best[0] >= INF:
        return None                                            # some step fits no compiled shape

    ssteps: List[SuperStep] = []
    k = 0
    while k < nr:
        cnt = take[k]
        pick = (cnt, pair[k] if cnt == 2 else single[k])
        cnt, (mi, perm) = pick
        bit = {ch: b for b, ch in enumerate(perm)}
        layers = []
        for g in raw[k:k + cnt]:
            st = sch.steps[g["step"]]
            f = factors[g["fi"]]
            pinned = [(j, bit[ch]) for j, ch in g["pinned"]]
            free = [(j, sum(1 << bit[ch] for ch in chs)) for j, chs in g["free"]]
            closed = [(sb, bit[ch]) for sb, ch in g["closed"]]
            pk = [sum(1 << bit[ch] for ch in chs) for chs in g["pk"]]
            NP, NF = len(pinned), len(free)
            T = np.zeros(1 << (NP + NF))
            for pidx in range(1 << NP):
                for kk in range(1 << NF):
                    a = 0
                    for q, (j, _) in enumerate(pinned):
                        a |= ((pidx >> q) & 1) << j
                    for q, (j, _) in enumerate(free):
                        a |= ((kk >> q) & 1) << j
                    T[(pidx << NF) | kk] = st.table[a]
            layers.append(Layer(g["step"], g["fi"], tuple(f.vars), pinned, free, closed, T, pk))
        ss = SuperStep(layers, list(perm), mi)
        ss.fresh = [list(g["fresh"]) for g in raw[k:k + cnt]]
        if cnt == 2:
            both = {pb for _, pb in layers[0].pinned} & {cb for _, cb in layers[1].closed}
            if both:
                b = both.pop()
                sb = [s_ for s_, cb in layers[1].closed if cb == b][0]
                ss.late = (sb, b)
                layers[1].closed = [(s_, cb) for s_, cb in layers[1].closed if cb != b]
        M = len(perm)
        ss.bpp = sum((1 << M) * len(l.free) for l in layers) if sch.semiring == S.MAXPLUS else 0
        ssteps.append(ss)
        k += cnt

    # positions: chains -> 0..W-1, shot bits -> W..NB-1
    chain_pos, n_conf = _assign_positions([ss.chains for ss in ssteps], W)
    # liveness of chains per super-step (a chain is dead after the step that closes it without re-opening)
    alive = set(range(W0))                                     # chains beyond W0 come alive when a fresh pin opens them
    wbase = 0
    for ss in ssteps:
        M = len(ss.chains)
        ss.pos = [chain_pos[ch] for ch in ss.chains]
        patch = set(ss.pos)
        nonpatch = [p for p in range(NB) if p not in patch]
        active = {chain_pos[ch] for ch in alive} | set(range(W, NB))
        lanes = []
        for r in range(4):
            cands = [p for p in nonpatch if p < 8 and p % 4 == r and p not in lanes]
            cands.sort(key=lambda p: (p not in active, p))
            if cands:
                lanes.append(cands[0])
        ss.conflict = len(lanes) < 4
        rest = [p for p in nonpatch if p not in lanes]
        rest.sort(key=lambda p: (p not in active, p >= W, p))      # active slot bits first, then shot bits, inactive last
        while len(lanes) < 5:
            lanes.append(rest.pop(0))
        ss.lanepos = lanes
        ss.looppos = [p for p in rest if p in active]
        n_iter = 1 << len(ss.looppos)
        if ss.bpp:
            ipw = 32 // ss.bpp
            ss.n_words = (n_iter + ipw - 1) // ipw
        ss.wbase = wbase
        wbase += ss.n_words
        # chains that come alive (fresh pins) and chains that die in this super-step, layer by layer
        for li, l in enumerate(ss.layers):
            alive.update(ss.fresh[li])
            reused = {pb for _, pb in l.pinned}
            for _, cb in l.closed:
                if cb not in reused:
                    alive.discard(ss.chains[cb])
        if ss.late is not None and ss.late[1] not in {pb for _, pb in ss.layers[1].pinned}:
            alive.discard(ss.chains[ss.late[1]])
    out_index = [0]
    if sch.semiring == S.SUMPROD:
        obs_pos = [0] * sch.n_obs
        for c in final_live:
            obs_pos[checks[c].index] = chain_pos[chain_of[c]]
        if sorted(chain_pos[chain_of[c]] for c in final_live) != sorted(chain_pos[ch] for ch in alive):
            return None
        out_index = [sum(((i >> o) & 1) << obs_pos[o] for o in range(sch.n_obs)) for i in range(1 << sch.n_obs)]
    elif alive:
        return None

    hs, hc = _head_eval(sch, h, roles, head_bits, live_order)
    if W > W0:
        # chains added by fresh pins are dead after the head: their set halves hold the semiring's zero
        ext = np.full((hs.shape[0], 1 << W), -np.inf if sch.semiring == S.MAXPLUS else 0.0)
        ext[:, : 1 << W0] = hs
        hs = ext
        if hc is not None:
            extc = np.zeros((hc.shape[0], 1 << W, hc.shape[2]), dtype=np.uint64)
            extc[:, : 1 << W0, :] = hc
            hc = extc
    # head table in POSITION order: index bit chain_pos[k] = parity of chain k
    idx = np.arange(1 << W)
    src = np.zeros_like(idx)
    for kch in range(W):
        src |= ((idx >> chain_pos[kch]) & 1) << kch
    hs = np.ascontiguousarray(hs[:, src])
    ncw = max(1, (sch.n_vars + 63) // 64)
    if hc is not None:
        hcw = np.ascontiguousarray(hc[:, src, :])
    else:
        hcw = np.zeros((hs.shape[0], 1, ncw), dtype=np.uint64)
    plan = SweepPlan(sch.semiring, sch.n_vars, sch.n_checks, sch.n_obs, W, sg, h, head_bits, hs, hcw, ssteps, wbase,
                     out_index, n_conf)
    _encode(plan)
    return plan


def _encode(p: SweepPlan):
    n = len(p.ssteps)
    rec = np.zeros((n, REC_INTS), dtype=np.int32)
    tb = np.zeros((n, TB_INTS), dtype=np.int32)
    lanetab = np.zeros((n, 32), dtype=np.uint32)
    tvals: List[float] = []
    submask = ((1 << p.sg) - 1) << p.W
    for i, ss in enumerate(p.ssteps):
        M = len(ss.pos)
        r = rec[i]
        r[0] = ss.menu
        r[1] = 1 << len(ss.looppos)
        while len(tvals) % 2:
            tvals.append(0.0)
        r[2] = len(tvals)
        for l in ss.layers:
            tvals += [float(x) for x in l.T]
        flipmask = 0
        if ss.late is not None:
            # second copy of layer 1's table with the rows of the pinned variable at the late bit swapped
            l = ss.layers[1]
            NF = len(l.free)
            for q, (_, pb) in enumerate(l.pinned):
                if pb == ss.late[1]:
                    flipmask |= 1 << q
            tvals += [float(l.T[((ix >> NF) ^ flipmask) << NF | (ix & ((1 << NF) - 1))]) for ix in range(len(l.T))]
        r[3] = ss.wbase
        ain = [phys(1 << ss.pos[b]) << 3 for b in range(M)] + [0] * (4 - M)
        r[4] = ain[0] | (ain[1] << 16)
        r[5] = ain[2] | (ain[3] << 16)
        r[6] = i * 32                                              # lane table row
        r[7] = i << p.sg                                           # row of the per-pass closed-bit table
        la, ls = [], []
        for it in range(8):
            x = 0
            for q, pos in enumerate(ss.looppos):
                x |= ((it >> q) & 1) << pos
            la.append(phys(x) << 3 if it < (1 << len(ss.looppos)) else 0)
            ls.append((x & submask) >> p.W if it < (1 << len(ss.looppos)) else 0)
        for q in range(4):
            r[8 + q] = la[2 * q] | (la[2 * q + 1] << 16)
        r[12] = ls[0] | (ls[1] << 8) | (ls[2] << 16) | (ls[3] << 24)
        r[13] = ls[4] | (ls[5] << 8) | (ls[6] << 16) | (ls[7] << 24)
        closed = [(sb, ss.pos[cb]) for l in ss.layers for sb, cb in l.closed]
        assert len(closed) <= 4
        r[14] = len(closed)
        for q, (sb, pos) in enumerate(closed):
            r[16 + q] = sb | (phys(1 << pos) << 19)
        r[20] = -1
        if ss.late is not None:
            r[20] = ss.late[0] | (phys(1 << ss.pos[ss.late[1]]) << 19) | (1 << 30 if flipmask else 0)
        for lane in range(32):
            x = 0
            for q, pos in enumerate(ss.lanepos):
                x |= ((lane >> q) & 1) << pos
            lanetab[i, lane] = (phys(x) << 3) | (((x & submask) >> p.W) << 16)
        # traceback record
        t = tb[i]
        t[0], t[1], t[2], t[3], t[4] = M, len(ss.layers), len(ss.looppos), ss.bpp, ss.wbase
        t[5] = (32 // ss.bpp) if ss.bpp else 0
        t[6] = len(closed)
        t[7] = -1 if ss.late is None else (ss.late[0] | (ss.late[1] << 16))
        for b in range(4):
            t[8 + b] = ss.pos[b] if b < M else -1
        for q in range(5):
            t[12 + q] = ss.lanepos[q]
            t[17 + q] = ss.looppos[q] if q < len(ss.looppos) else -1
        for q, (sb, pos) in enumerate(closed):
            t[22 + 2 * q], t[23 + 2 * q] = sb, pos
        bpoff = 0
        for li, l in enumerate(ss.layers):
            o = 30 + 14 * li
            t[o], t[o + 1], t[o + 2] = len(l.pinned), len(l.free), bpoff
            for q in range(2):
                if q < len(l.pinned):
                    t[o + 3 + 2 * q], t[o + 4 + 2 * q] = l.pinned[q][1], l.vars[l.pinned[q][0]]
                else:
                    t[o + 3 + 2 * q], t[o + 4 + 2 * q] = -1, -1
                if q < len(l.free):
                    t[o + 7 + 2 * q], t[o + 8 + 2 * q] = l.free[q][1], l.vars[l.free[q][0]]
                else:
                    t[o + 7 + 2 * q], t[o + 8 + 2 * q] = 0, -1
            t[o + 11] = flipmask if (li == 1 and ss.late is not None) else 0
            for q in range(len(l.pinned)):
                t[o + 12 + q] = l.pk[q] if q < len(l.pk) else 0
            bpoff += (1 << M) * len(l.free)
    p.rec, p.tb, p.lanetab = rec, tb, lanetab
    p.tvals = np.asarray(tvals if tvals else [0.0], dtype=np.float64)
