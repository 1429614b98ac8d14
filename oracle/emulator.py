"""Table-level emulator of the lowered schedule (oracle; see oracle/__init__.py).

Executes the flat arrays `hdr / ints / tables` emitted by `tensorqec.jl_b200.schedule._encode` with exactly the
index arithmetic of the CUDA kernels (include/tqec.h, csrc/tqec_decode.cu): re-insert closed bits, read the opened
pattern, gather over the coset, strict-greater update, 1 back-pointer of kb bits per output, serial traceback.
Output bit b of a step lives at full slot perm[b]; the closed slots carry the shot's syndrome bits.
It exists to check the LOWERING on a CPU-only box; the recurrence itself is checked by frontier.py.
"""
import numpy as np

(H_R, H_WIN, H_NOPEN, H_NCLOSE, H_WOUT, H_NK, H_KB, H_OFF_T, H_OFF_ML, H_OFF_MK, H_OFF_A0, H_OFF_KER, H_OFF_VARS,
 H_OFF_CLOSE) = range(14)


def _scatter(sch, h, tau, syn, w_out, n_close):
    """output index -> full index: bit b of tau goes to slot perm[b]; closed slots take the shot's syndrome bits.
    The closed list (slot, syndrome bit) sits at ints[off_close ...], perm[w_out] right behind it."""
    off = int(h[H_OFF_CLOSE])
    perm = sch.ints[off + 2 * n_close: off + 2 * n_close + w_out]
    full = np.zeros_like(tau)
    for b in range(w_out):
        full |= ((tau >> b) & 1) << int(perm[b])
    for c in range(n_close):
        slot, bit = int(sch.ints[off + 2 * c]), int(sch.ints[off + 2 * c + 1])
        full |= syn[:, bit][:, None] << slot
    return full


def run(sch, syndromes):
    """sch: tensorqec.jl_b200.schedule.Schedule; syndromes (B, n_checks) 0/1.
    Max-plus -> (logp, config (B, n_vars));  sum-product -> marginal (B, 2^n_obs) (observable 0 fastest)."""
    syn = np.atleast_2d(np.asarray(syndromes, dtype=np.int64))
    B = syn.shape[0]
    maxplus = sch.semiring == 0
    S = np.full((B, 1), 0.0 if maxplus else 1.0)
    bps = []
    for h in sch.hdr:
        w_in, n_open, n_close, w_out, nk = (int(h[i]) for i in (H_WIN, H_NOPEN, H_NCLOSE, H_WOUT, H_NK))
        tau = np.arange(1 << w_out, dtype=np.int64)[None, :].repeat(B, axis=0)
        full = _scatter(sch, h, tau, syn, w_out, n_close)
        pat = full >> w_in
        low = (full & ((1 << w_in) - 1)) ^ sch.ints[h[H_OFF_ML] + pat]
        best = None
        bk = np.zeros_like(tau)
        for k in range(nk):
            src = low ^ int(sch.ints[h[H_OFF_MK] + k])
            tv = sch.tables[h[H_OFF_T] + pat * nk + k]
            sv = np.take_along_axis(S, src, axis=1)
            v = sv + tv if maxplus else sv * tv
            if best is None:
                best = v
            elif maxplus:
                upd = v > best
                best = np.where(upd, v, best)
                bk = np.where(upd, k, bk)
            else:
                best = best + v
        S = best
        bps.append(bk)
    if not maxplus:
        out = np.zeros((B, 1 << sch.n_obs))
        idx = np.arange(1 << sch.n_obs)
        src = np.zeros_like(idx)
        for i, s in enumerate(sch.obs_slot):
            src |= ((idx >> i) & 1) << s
        return np.ldexp(S[:, src], getattr(sch, 'log2_scale', 0))
    logp = S[:, 0]
    config = np.zeros((B, sch.n_vars), dtype=np.uint8)
    tau = np.zeros(B, dtype=np.int64)
    for t in range(len(sch.hdr) - 1, -1, -1):
        h = sch.hdr[t]
        w_in, n_open, n_close, w_out, nk, r = (int(h[i]) for i in (H_WIN, H_NOPEN, H_NCLOSE, H_WOUT, H_NK, H_R))
        k = bps[t][np.arange(B), tau]
        full = _scatter(sch, h, tau[:, None], syn, w_out, n_close)[:, 0]
        pat = full >> w_in
        a = sch.ints[h[H_OFF_A0] + pat] ^ sch.ints[h[H_OFF_KER] + k]
        for j in range(r):
            config[:, int(sch.ints[h[H_OFF_VARS] + j])] = (a >> j) & 1
        tau = (full & ((1 << w_in) - 1)) ^ sch.ints[h[H_OFF_ML] + pat] ^ sch.ints[h[H_OFF_MK] + k]
    return logp, config
