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


def _bf_pass(St, bf, vals, syn_b, t_in, t_out):
    """One butterfly pass (k_wide_bf, encoding of tqec_lower_wide.cpp:bf_encode_pass) on the tiles `St` (tiles, 2^t_in)
    of one shot: dead entries are zero-filled at the load, a coset is loaded, zeroed where the group says so, run through
    the group's unit steps (ratio 0 = padding) and dependent steps, stored; closed checks only update the address mask."""
    n_groups, n_dep, n_closes, _, n_pos, bt_in, bt_out, G = (int(bf[k]) for k in range(8))
    assert bt_in == t_in and bt_out == t_out
    pout = [int(x) for x in bf[8:20]]
    grec = bf[20:20 + 16 * n_groups].reshape(n_groups, 16)
    drec = bf[20 + 16 * n_groups:][:n_dep]
    crec = bf[20 + 16 * n_groups + n_dep:][:2 * n_closes].reshape(n_closes, 2)
    S = np.zeros((St.shape[0], 1 << n_pos))
    S[:, : 1 << t_in] = St
    X = 0
    NE = 1 << G
    seen = np.zeros(1 << n_pos, dtype=bool)

    def pair(v, c, r):
        low = c & -c
        nv = v.copy()
        for k in range(NE):
            if not (k & low):
                a, b = v[:, :, k], v[:, :, k ^ c]
                nv[:, :, k] = r * b + a
                nv[:, :, k ^ c] = r * a + b
        return nv

    for g in grec:
        n_free, dep0, nd, close0, nc = (int(g[k]) for k in range(5))
        basis = [int(g[5 + j]) for j in range(G)]
        Z, val0 = int(g[10]) & 0xFFFFFFFF, int(g[11])
        order = [(int(g[13]) >> (4 * i)) & 15 for i in range(n_free)]
        i = np.arange(1 << n_free, dtype=np.int64)
        rep = np.zeros_like(i)
        for bpos, q in enumerate(order):
            rep |= ((i >> bpos) & 1) << q
        cm = np.zeros(NE, dtype=np.int64)
        for k in range(NE):
            for j in range(G):
                if (k >> j) & 1:
                    cm[k] ^= basis[j]
        idx = (rep[:, None] ^ cm[None, :]) ^ X                                  # (cosets, 2^G) actual indices
        assert np.unique(idx).size == idx.size, "cosets overlap"
        v = S[:, idx]                                                           # (tiles, cosets, 2^G)
        for k in range(NE):
            if (Z >> k) & 1:
                v[:, :, k] = 0.0
        for j in range(G):
            v = pair(v, 1 << j, float(vals[val0 + j]))
        for s in range(nd):
            c = int(drec[dep0 + s])
            assert 0 < c < NE
            v = pair(v, c, float(vals[val0 + G + s]))
        S[:, idx] = v
        for cl in crec[close0:close0 + nc]:
            if syn_b[int(cl[1])]:
                X ^= 1 << int(cl[0])
    l = np.arange(1 << t_out, dtype=np.int64)
    logical = np.zeros_like(l)
    for bpos in range(t_out):
        logical |= ((l >> bpos) & 1) << pout[bpos]
    return S[:, logical ^ X]


def run_tables(ph, sh, ints, tabs, obs_pos, n_obs, log2_scale, syndromes, bf=None):
    """Execute flat tables: `ph` (n_pass, 16), `sh` (n_steps, 16), pools; bf = None or (bf_off, bf_ints, bf_vals, mant, log2)
    to run the passes that have a butterfly block the way k_wide_bf does."""
    syn = np.atleast_2d(np.asarray(syndromes, dtype=np.uint8))
    B = syn.shape[0]
    out = np.zeros((B, 1 << n_obs))
    for b in range(B):
        G = np.ones(1)
        for ip, h in enumerate(ph):
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
            if bf is not None and int(bf[0][ip]) >= 0:
                St = _bf_pass(St, bf[1][int(bf[0][ip]):], bf[2], syn[b], t_in, t_out)
            else:
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
        idx = np.arange(1 << n_obs, dtype=np.int64)
        src = np.zeros_like(idx)
        for o in range(n_obs):
            src |= ((idx >> o) & 1) << int(obs_pos[o])
        out[b] = G[src]
    if bf is not None:
        return np.ldexp(out * float(bf[3]), int(log2_scale) + int(bf[4]))
    return np.ldexp(out, int(log2_scale))


def run(plan, syndromes):
    """syndromes (B, n_checks) 0/1 -> marginals (B, 2^n_obs), observable 0 fastest, static scaling undone."""
    return run_tables(plan.pass_hdr, plan.step_hdr, plan.ints, plan.tables, plan.obs_pos, plan.n_obs, plan.log2_scale, syndromes)


def run_lowered(lw, n_obs, syndromes, butterfly=True):
    """The same on the tables of the library's own lowering (`_cabi.Lowered`), butterfly passes included."""
    from tensorqec.jl_b200 import _cabi as A
    m = lw.meta
    ph = lw.get(A.LW_WD_PASS_HDR).reshape(-1, W.PASS_INTS)
    sh = lw.get(A.LW_WD_STEP_HDR).reshape(-1, W.STEP_INTS)
    bf = None
    if butterfly:
        sc = lw.get(A.LW_WD_BF_SCALE)
        bf = (lw.get(A.LW_WD_BF_OFF), lw.get(A.LW_WD_BF_INTS), lw.get(A.LW_WD_BF_VALS), float(sc[0]), int(sc[1]))
    return run_tables(ph, sh, lw.get(A.LW_WD_INTS), lw.get(A.LW_WD_TABLES), lw.get(A.LW_WD_OBS_POS), n_obs, m["log2_scale"],
                      syndromes, bf)
