"""Table-level emulator of the in-place patch sweep (oracle; see oracle/__init__.py).

Executes the flat tables emitted by `tensorqec.jl_b200.sweep._encode` (rec / tb / lanetab / tvals, head tables) with
exactly the data movement of `k_sweep` (csrc/tqec_sweep.cu): head-table copy, per-step closed-bit address folding,
patch loads through the lane / loop / patch XOR masks, in-register layers, in-place stores, packed back-pointer words,
per-shot traceback over the traceback records.  It checks the LOWERING on a CPU-only box; the recurrence itself is
checked against frontier.py (tests/test_sweep_cpu.py).
"""
import numpy as np

MAXPLUS = 0


def _layers_of(menu_entry):
    M, layers = menu_entry
    return M, [(list(l[0]), list(l[1]), list(l[2]) if len(l) > 2 else [0] * len(l[0])) for l in layers]


def run(plan, menu, syndromes):
    """plan: tensorqec.jl_b200.sweep.SweepPlan; syndromes (B, n_checks) 0/1.
    Max-plus -> (logp (B,), config (B, n_vars) uint8); sum-product -> marginal (B, 2^n_obs)."""
    syn = np.atleast_2d(np.asarray(syndromes, dtype=np.int64))
    B = syn.shape[0]
    maxplus = plan.semiring == MAXPLUS
    W, sg = plan.W, plan.sg
    SG = 1 << sg
    NS = 1 << (W + sg)
    phys = lambda x: x ^ ((x >> 4) & 15)
    n_pass = (B + SG - 1) // SG
    logp = np.zeros(B)
    cfg_out = np.zeros((B, plan.n_vars), dtype=np.uint8)
    mar = np.zeros((B, 1 << plan.n_obs))
    n_ss = len(plan.rec)
    for ps in range(n_pass):
        shots = [min(ps * SG + s, B - 1) for s in range(SG)]
        ssyn = syn[shots]                                        # (SG, n_checks)
        hp = np.zeros(SG, dtype=np.int64)
        for j, b in enumerate(plan.head_bits):
            hp |= ssyn[:, b] << j
        state = np.full(NS, np.nan)
        for sub in range(SG):
            for e in range(1 << W):
                state[phys(e | (sub << W))] = plan.head_state[hp[sub], e]
        stab = np.zeros((n_ss, SG), dtype=np.int64)
        for i in range(n_ss):
            r = plan.rec[i]
            for q in range(int(r[14])):
                sb, pm = int(r[16 + q]) & 0xFFFF, (int(r[16 + q]) >> 16) & 0xFFFF
                stab[i] ^= np.where(ssyn[:, sb] == 1, pm, 0)
        late = np.zeros((n_ss, SG), dtype=np.int64)               # byte mask | 0x8000 when the late syndrome bit is set
        for i in range(n_ss):
            v = int(plan.rec[i][20])
            if v >= 0:
                sb, pm, fl = v & 0xFFFF, (v >> 16) & 0x3FFF, (v >> 30) & 1
                late[i] = np.where(ssyn[:, sb] == 1, pm | (0x8000 if fl else 0x4000), 0)
        bp = np.zeros((max(plan.bp_words, 1), 32), dtype=np.uint64)
        for i in range(n_ss):
            r = [int(v) for v in plan.rec[i]]
            M, layers = _layers_of(menu[r[0]])
            n_iter, toff, wbase = r[1], r[2], r[3]
            ain = [r[4] & 0xFFFF, (r[4] >> 16) & 0xFFFF, r[5] & 0xFFFF, (r[5] >> 16) & 0xFFFF]
            la = [(r[8 + q // 2] >> (16 * (q % 2))) & 0xFFFF for q in range(8)]
            ls = [(r[12 + q // 4] >> (8 * (q % 4))) & 0xFF for q in range(8)]
            pa = [0] * (1 << M)
            for j in range(1 << M):
                for b in range(M):
                    if (j >> b) & 1:
                        pa[j] ^= ain[b]
            bpp = sum((1 << M) * len(fm) for _, fm, _ in layers) if maxplus else 0
            ipw = 32 // bpp if bpp else 1
            for lane in range(32):
                lt = int(plan.lanetab[i, lane])
                laddr, lsub = lt & 0xFFFF, lt >> 16
                word = 0
                for it in range(n_iter):
                    base = laddr ^ la[it]
                    sub = lsub | ls[it]
                    inb = base ^ int(stab[i, sub])
                    lt_ = int(late[i, sub])
                    outb = base ^ (lt_ & 0x3FFF)
                    R = [state[(inb ^ pa[j]) >> 3] for j in range(1 << M)]
                    bits = 0
                    off = 0
                    to = toff
                    for li, (pb, fm, pk) in enumerate(layers):
                        NP, NF = len(pb), len(fm)
                        if li == 1 and (lt_ & 0x8000):
                            to += 1 << (NP + NF)                 # the row-swapped copy of layer 1's table
                        out = [0.0] * (1 << M)
                        for j in range(1 << M):
                            pidx = sum(((j >> pb[q]) & 1) << q for q in range(NP))
                            best, bk = None, 0
                            pflip = 0
                            for q in range(NP):
                                if (j >> pb[q]) & 1:
                                    pflip ^= pk[q]
                            for k in range(1 << NF):
                                src = j ^ pflip
                                for f in range(NF):
                                    if (k >> f) & 1:
                                        src ^= fm[f]
                                tv = plan.tvals[to + (pidx << NF) + k]
                                v = R[src] + tv if maxplus else R[src] * tv
                                if best is None:
                                    best = v
                                elif maxplus:
                                    if v > best:
                                        best, bk = v, k
                                else:
                                    best = best + v
                            out[j] = best
                            bits |= bk << (off + j * NF)
                        R = out
                        off += (1 << M) * NF
                        to += 1 << (NP + NF)
                    for j in range(1 << M):
                        state[(outb ^ pa[j]) >> 3] = R[j]
                    if bpp:
                        word |= bits << (bpp * (it % ipw))
                        if it % ipw == ipw - 1 or it == n_iter - 1:
                            bp[wbase + it // ipw, lane] = word
                            word = 0
        for sub in range(SG):
            shot = ps * SG + sub
            if shot >= B:
                continue
            if not maxplus:
                for idx, oi in enumerate(plan.out_index):
                    mar[shot, idx] = state[phys(oi | (sub << W))]
                continue
            x = plan.out_index[0] | (sub << W)
            logp[shot] = state[phys(x)]
            for i in range(n_ss - 1, -1, -1):
                t = [int(v) for v in plan.tb[i]]
                M, nl, nlb, bpp, wbase, ipw, ncl = t[0:7]
                pos = t[8:8 + M]
                j = sum(((x >> pos[b]) & 1) << b for b in range(M))
                lane = sum(((x >> t[12 + q]) & 1) << q for q in range(5))
                it = sum(((x >> t[17 + q]) & 1) << q for q in range(nlb))
                lsyn = 0
                if t[7] >= 0:
                    lsyn = int(ssyn[sub, t[7] & 0xFFFF])
                    j ^= lsyn << (t[7] >> 16)
                pbits = 0
                if bpp:
                    pbits = (int(bp[wbase + it // ipw, lane]) >> (bpp * (it % ipw))) & ((1 << bpp) - 1)
                for li in range(nl - 1, -1, -1):
                    o = 30 + 14 * li
                    NP, NF, bpoff = t[o], t[o + 1], t[o + 2]
                    k = (pbits >> (bpoff + j * NF)) & ((1 << NF) - 1) if NF else 0
                    pflip = 0
                    for q in range(NP):
                        cfg_out[shot, t[o + 4 + 2 * q]] = ((j >> t[o + 3 + 2 * q]) & 1) ^ (lsyn & (t[o + 11] >> q) & 1)
                        if (j >> t[o + 3 + 2 * q]) & 1:
                            pflip ^= t[o + 12 + q]
                    j ^= pflip
                    for q in range(NF):
                        cfg_out[shot, t[o + 8 + 2 * q]] = (k >> q) & 1
                        if (k >> q) & 1:
                            j ^= t[o + 7 + 2 * q]
                for b in range(M):
                    x &= ~(1 << pos[b])
                    x |= ((j >> b) & 1) << pos[b]
                for q in range(ncl):
                    if ssyn[sub, t[22 + 2 * q]]:
                        x ^= 1 << t[23 + 2 * q]
            e = x & ((1 << W) - 1)
            words = plan.head_cfg[hp[sub], e]
            for v in range(plan.n_vars):
                if (int(words[v >> 6]) >> (v & 63)) & 1:
                    cfg_out[shot, v] = 1
    if maxplus:
        return logp, cfg_out
    return mar
