"""Table-level emulator of the lowered schedule (oracle; see oracle/__init__.py).

Executes the flat arrays `hdr / ints / tables` emitted by `tensorqec.jl_b200.schedule._encode` with exactly the
index arithmetic of the CUDA kernels (include/tqec.h, csrc/tqec_decode.cu): re-insert closed bits, read the opened
pattern, gather over the coset, strict-greater update, 1 back-pointer of kb bits per output, serial traceback.
It exists to check the LOWERING on a CPU-only box; the recurrence itself is checked by frontier.py.
"""
import numpy as np

(H_R, H_WIN, H_NOPEN, H_NCLOSE, H_WOUT, H_NK, H_KB, H_OFF_T, H_OFF_ML, H_OFF_MK, H_OFF_A0, H_OFF_KER, H_OFF_VARS,
 H_OFF_CLOSE) = range(14)


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
        full = tau.copy()
        for c in range(n_close):
            slot = int(sch.ints[h[H_OFF_CLOSE] + 2 * c])
            bit = int(sch.ints[h[H_OFF_CLOSE] + 2 * c + 1])
            low = full & ((1 << slot) - 1)
            full = ((full >> slot) << (slot + 1)) | (syn[:, bit][:, None] << slot) | low
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
        return S[:, src]
    logp = S[:, 0]
    config = np.zeros((B, sch.n_vars), dtype=np.uint8)
    tau = np.zeros(B, dtype=np.int64)
    for t in range(len(sch.hdr) - 1, -1, -1):
        h = sch.hdr[t]
        w_in, n_open, n_close, w_out, nk, r = (int(h[i]) for i in (H_WIN, H_NOPEN, H_NCLOSE, H_WOUT, H_NK, H_R))
        k = bps[t][np.arange(B), tau]
        full = tau.copy()
        for c in range(n_close):
            slot = int(sch.ints[h[H_OFF_CLOSE] + 2 * c])
            bit = int(sch.ints[h[H_OFF_CLOSE] + 2 * c + 1])
            low = full & ((1 << slot) - 1)
            full = ((full >> slot) << (slot + 1)) | (syn[:, bit] << slot) | low
        pat = full >> w_in
        a = sch.ints[h[H_OFF_A0] + pat] ^ sch.ints[h[H_OFF_KER] + k]
        for j in range(r):
            config[:, int(sch.ints[h[H_OFF_VARS] + j])] = (a >> j) & 1
        tau = (full & ((1 << w_in) - 1)) ^ sch.ints[h[H_OFF_ML] + pat] ^ sch.ints[h[H_OFF_MK] + k]
    return logp, config
