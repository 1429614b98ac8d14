"""Table-level emulator of the global-memory executor (oracle; see oracle/__init__.py): executes the FLAT tables of a
`WidePlan` (tensorqec.jl_b200/wide.py) exactly as `k_wide_pass` reads them -- pass header, tile gather by bit deposit,
local steps by scatter look-up + gather, tile store -- vectorised over tiles with numpy.  Used by the CPU tests to check
the lowering against the recurrence oracle (frontier.py) before any GPU time is spent."""
import numpy as np

from tensorqec.jl_b200 import wide as W


def _pdep(x, mask):
    """deposit the low bits of x (array) at the set bits of mask."""
    out = np.zeros_like(x)
    k = 0
    b = 0
    while mask >> b:
        if (mask >> b) & 1:
            out |= ((x >> k) & 1) << b
            k += 1
        b += 1
    return out


def run(plan, syndromes):
    """syndromes (B, n_checks) 0/1 -> marginals (B, 2^n_obs), observable 0 fastest, static scaling undone."""
    syn = np.atleast_2d(np.asarray(syndromes, dtype=np.uint8))
    B = syn.shape[0]
    out = np.zeros((B, 1 << plan.n_obs))
    ph, sh, ints, tabs = plan.pass_hdr, plan.step_hdr, plan.ints, plan.tables
    for b in range(B):
        G = np.ones(1)
        for h in ph:
            w_in, w_out, t_in, t_out, ns, s0 = (int(h[k]) for k in (W.P_WIN, W.P_WOUT, W.P_TIN, W.P_TOUT, W.P_NSTEPS, W.P_STEP0))
            tin, tout = int(h[W.P_TINMASK]), int(h[W.P_TOUTMASK])
            I = ints[int(h[W.P_OFF_INTS]): int(h[W.P_OFF_INTS]) + int(h[W.P_N_INTS])]
            T = tabs[int(h[W.P_OFF_TAB]): int(h[W.P_OFF_TAB]) + int(h[W.P_N_TAB])]
            assert G.size == 1 << w_in
            n_spec = w_in - t_in
            assert n_spec == w_out - t_out
            sp = np.arange(1 << n_spec, dtype=np.int64)
            base_in = _pdep(sp, ((1 << w_in) - 1) & ~tin)
            base_out = _pdep(sp, ((1 << w_out) - 1) & ~tout)
            loc = np.arange(1 << t_in, dtype=np.int64)
            St = G[base_in[:, None] | _pdep(loc, tin)[None, :]]                 # (tiles, 2^t_in)
            for q in sh[s0:s0 + ns]:
                lw_in, n_open, n_close, lw_out, nk = (int(q[k]) for k in (W.L_WIN, W.L_NOPEN, W.L_NCLOSE, W.L_WOUT, W.L_NK))
                ML = I[int(q[W.L_OFF_ML]):][: 1 << n_open].astype(np.int64)
                MK = I[int(q[W.L_OFF_MK]):][:nk].astype(np.int64)
                CL = I[int(q[W.L_OFF_CLOSE]):][: 2 * n_close]
                Tt = T[int(q[W.L_OFF_T]):][: (1 << n_open) * nk]
                assert St.shape[1] == 1 << lw_in
                cb = 0
                for c in range(n_close):
                    cb |= int(syn[b, CL[2 * c + 1]]) << int(CL[2 * c])
                tau = np.arange(1 << lw_out, dtype=np.int64)
                full = _pdep(tau, int(q[W.L_KEEPMASK])) | cb
                pat = full >> lw_in
                low = (full & ((1 << lw_in) - 1)) ^ ML[pat]
                acc = St[:, low ^ MK[0]] * Tt[pat * nk]
                for k in range(1, nk):
                    acc = acc + St[:, low ^ MK[k]] * Tt[pat * nk + k]
                St = acc
            assert St.shape[1] == 1 << t_out
            Gn = np.zeros(1 << w_out)
            Gn[base_out[:, None] | _pdep(np.arange(1 << t_out, dtype=np.int64), tout)[None, :]] = St
            G = Gn
        idx = np.arange(1 << plan.n_obs, dtype=np.int64)
        src = np.zeros_like(idx)
        for o in range(plan.n_obs):
            src |= ((idx >> o) & 1) << plan.obs_pos[o]
        out[b] = G[src]
    return np.ldexp(out, plan.log2_scale)
