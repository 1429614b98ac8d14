"""The frontier recurrence the CUDA kernels execute, written independently (oracle; see oracle/__init__.py).

Input is the factor graph and the absorption ORDER only (no slot layout, no lowered tables): the state is a numpy
array with one size-2 axis per currently open check, keyed by check id, batched over shots.  Per step (factor f with
variables v_1..v_r, table T[a], a = sum_j a_j 2^j):
    candidates a in ascending order (or in the schedule's enumeration order, `priority_of`: the tie rule is "first best
    candidate in enumeration order");  cand_a[sigma'] = S[sigma' xor M(a)] (x) T[a], where M(a) flips
    the axis of every touched check whose variables have odd parity under a; checks opened by the step enter with
    parity 0; the result keeps the FIRST best candidate (strict >, i.e. the smallest a on exact FP64 ties); checks
    whose last variable was absorbed are then indexed at the shot's syndrome bit and dropped.
This is a pairwise contraction of the XOR-factorised network of tndecoder.jl:221-238 along a caterpillar tree; each
candidate is one IEEE add (max-plus) so forward values are bit-reproducible by any implementation of the recurrence.
"""
import math

import numpy as np

MAXPLUS, SUMPROD = 0, 1


def _parity_flips(factor_vars, checks, touched, a):
    flips = []
    for c in touched:
        p = 0
        for j, v in enumerate(factor_vars):
            if v in checks[c].vars:
                p ^= (a >> j) & 1
        if p:
            flips.append(c)
    return flips


def priority_of(sch):
    """Candidate enumeration order of a lowered schedule, per step: candidate k of opened pattern p is the assignment
    a0[p] ^ ker[k]; within a pattern the kernels compare k = 0, 1, .. and keep the first best one."""
    out = []
    for st in sch.steps:
        out.append([int(st.a0[p]) ^ int(k) for k in st.ker for p in range(len(st.a0)) if st.a0[p] >= 0])
    return out


def run(factors, checks, order, semiring, syndromes, n_vars, want_config=True, priority=None, rescale=False):
    """factors[i].vars/.table (flat, first variable fastest), checks[c].vars/.kind/.index.
    syndromes: (B, n_syn) 0/1.  Max-plus -> (logp (B,), config (B, n_vars) uint8);  sum-product -> marginal
    (B, 2^n_obs) with sector index = sum_i obs_i << i."""
    syndromes = np.atleast_2d(np.asarray(syndromes, dtype=np.uint8))
    B = syndromes.shape[0]
    maxplus = semiring == MAXPLUS
    zero = -np.inf if maxplus else 0.0
    owner = {v: i for i, f in enumerate(factors) for v in f.vars}
    c_factors = [sorted({owner[v] for v in c.vars}) for c in checks]
    remaining = [len(x) for x in c_factors]
    axes = []                                                    # check ids, axis k+1 of S <-> axes[k]
    S = np.full((B,), 0.0 if maxplus else 1.0)
    exps = np.zeros(B, dtype=np.int64)
    trace = []
    orphan = [ci for ci, fs in enumerate(c_factors) if not fs]
    for t, fi in enumerate(order):
        f = factors[fi]
        # log-weights through libm's scalar log (numpy's SIMD log differs from it in the last bit on ~0.3 % of arguments;
        # the recurrence is defined on the libm values, which is also what the product's lowerings use)
        T = (np.array([math.log(x) if x > 0.0 else -math.inf for x in np.asarray(f.table, dtype=np.float64).reshape(-1)])
             if maxplus else np.asarray(f.table, dtype=np.float64))
        touched = [c for c in range(len(checks)) if fi in c_factors[c]]
        opened = [c for c in touched if c not in axes]
        if t == 0:
            opened = opened + orphan
        for c in opened:                                         # new axis, mass only at parity 0
            S = np.stack([S, np.full_like(S, zero)], axis=-1)
            axes.append(c)
        best = None
        arg = None
        cand_list = range(1 << len(f.vars)) if priority is None else priority[t]
        for a in cand_list:
            flips = _parity_flips(f.vars, checks, touched, a)
            src = np.flip(S, axis=tuple(axes.index(c) + 1 for c in flips)) if flips else S
            cand = src + T[a] if maxplus else src * T[a]
            if best is None:
                best = cand.copy()
                arg = np.zeros(S.shape, dtype=np.int16)
            elif maxplus:
                upd = cand > best
                best = np.where(upd, cand, best)
                arg = np.where(upd, np.int16(a), arg)
            else:
                best = best + cand
        S = best
        if rescale and not maxplus:
            # dynamic rescaling, per shot: pull the largest entry back to [1, 2) and book the exponent
            mx = S.reshape(B, -1).max(axis=1)
            e = np.where(mx > 0, np.floor(np.log2(np.where(mx > 0, mx, 1.0))), 0).astype(np.int64)
            S = np.ldexp(S, (-e).reshape([B] + [1] * (S.ndim - 1)).astype(np.int32))
            exps = exps + e
        closing = []
        for c in touched + (orphan if t == 0 else []):
            if c in orphan:
                if checks[c].kind == "syn":
                    closing.append(c)
                continue
            remaining[c] -= 1
            if remaining[c] == 0 and checks[c].kind == "syn":
                closing.append(c)
        full_axes = list(axes)
        for c in closing:                                        # clamp to the syndrome bit and drop the axis
            k = axes.index(c) + 1
            bit = syndromes[:, checks[c].index].astype(np.intp)
            shp = [B] + [1] * (S.ndim - 1)
            ix = bit.reshape(shp)
            S = np.take_along_axis(S, ix, axis=k).squeeze(axis=k)
            if maxplus:
                arg = np.take_along_axis(arg, ix, axis=k).squeeze(axis=k)
            axes.pop(k - 1)
        if maxplus:
            trace.append((fi, touched, full_axes, list(axes), closing, arg))
    if not maxplus:
        # remaining axes are observables; order them by observable index, first observable fastest
        obs_order = sorted(range(len(axes)), key=lambda k: checks[axes[k]].index)
        S = np.transpose(S, [0] + [k + 1 for k in obs_order])
        S = S.reshape(B, -1, order="F") if S.ndim > 1 else S.reshape(B, 1)
        return (S, exps) if rescale else S
    assert S.ndim == 1
    logp = S
    if not want_config:
        return logp, None
    config = np.zeros((B, n_vars), dtype=np.uint8)
    cur = {}                                                     # check id -> (B,) parity bits of the current index
    for fi, touched, full_axes, out_axes, closing, arg in reversed(trace):
        f = factors[fi]
        for c in closing:
            cur[c] = syndromes[:, checks[c].index].copy()
        idx = (np.arange(B),) + tuple(cur[c].astype(np.intp) for c in out_axes)
        a = arg[idx]
        for j, v in enumerate(f.vars):
            config[:, v] = (a >> j) & 1
        for c in touched:
            p = np.zeros(B, dtype=np.uint8)
            for j, v in enumerate(f.vars):
                if v in checks[c].vars:
                    p ^= ((a >> j) & 1).astype(np.uint8)
            cur[c] = cur[c] ^ p
        opened_here = [c for c in full_axes if c not in _prev_axes(trace, fi)]
        for c in opened_here:
            cur.pop(c, None)
    return logp, config


def _prev_axes(trace, fi):
    """axes alive before the step that absorbed factor fi."""
    for k, rec in enumerate(trace):
        if rec[0] == fi:
            return trace[k - 1][3] if k > 0 else []
    raise KeyError(fi)
