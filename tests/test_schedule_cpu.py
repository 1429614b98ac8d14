"""Host logic on CPU: the lowering (schedule.py) against the independent recurrence, the dense reference network and
exhaustive enumeration; the C ABI's export list; loud failure without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import bruteforce, cref, dense, emulator, frontier, gf2, networks, philox

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _random_syndromes(t, em, seed, B):
    ex, ez = philox.sample_depolarizing(em.px, em.py, em.pz, seed, 0, B)
    sx, sz = gf2.css_syndrome(ex, ez, t.stgx.H, t.stgz.H)
    return np.concatenate([sx, sz], axis=1)


def test_tnmap_d3_exhaustive_all_oracles_agree(tq):
    t = tq.CSSTannerGraph(tq.SurfaceCode(3, 3))
    em = tq.iid_error(0.05, t)
    gdp, _ = tq.reduce2general(t, em)
    sch = tq.tnmap_schedule(tq.TNMAP(), gdp)
    syn = ((np.arange(256)[:, None] >> np.arange(8)) & 1).astype(np.uint8)
    lp_f, cfg_f = frontier.run(sch.factors, sch.checks, sch.order, 0, syn, 18, priority=frontier.priority_of(sch))
    lp_e, cfg_e = emulator.run(sch, syn)
    lp_c, cfg_c = cref.FrontierPlan(sch).run(syn)
    assert np.array_equal(lp_f, lp_e) and np.array_equal(cfg_f, cfg_e)
    assert np.array_equal(lp_f, lp_c) and np.array_equal(cfg_f, cfg_c)
    nq, s2q, pix, pri = networks.general_problem_css(t, em.px, em.py, em.pz)
    en = bruteforce.Enumeration(nq, s2q, pix, pri)
    dp = cref.DensePlan(networks.tnmap_network(nq, s2q, pix, pri), 8, 18, True)
    lp_d, cfg_d = dp.run(syn)
    n_unique = 0
    for b in range(256):
        best, maxs = en.map(syn[b])
        assert abs(lp_f[b] - best) <= 1e-12 * abs(best) and abs(lp_d[b] - best) <= 1e-12 * abs(best)
        assert any(np.array_equal(cfg_f[b], m) for m in maxs)
        assert any(np.array_equal(cfg_d[b], m) for m in maxs)
        if len(maxs) == 1:
            n_unique += 1
            assert np.array_equal(cfg_f[b], cfg_d[b])           # identical except on ties
    assert 0 < n_unique < 256                                   # ties are common with (p, p, p) noise
    # numpy dense executor == C dense executor on a few syndromes
    for b in (0, 37, 255):
        lp, cfg = dense.most_probable_config(networks.tnmap_network(nq, s2q, pix, pri, syn[b]))
        assert abs(lp - lp_d[b]) < 1e-12 and np.array_equal(cfg[:18], cfg_d[b])


@pytest.mark.parametrize("d", [5, 7])
def test_tnmap_lowering_vs_recurrence_and_dense(tq, d):
    t = tq.CSSTannerGraph(tq.SurfaceCode(d, d))
    rng = np.random.default_rng(d)
    n = d * d
    em = tq.IndependentDepolarizingError(rng.uniform(0.01, 0.1, n), rng.uniform(0.01, 0.1, n), rng.uniform(0.01, 0.1, n))
    gdp, _ = tq.reduce2general(t, em)
    sch = tq.tnmap_schedule(tq.TNMAP(), gdp)
    assert sch.w_max <= d + 2
    syn = _random_syndromes(t, em, d, 48)
    lp_f, cfg_f = frontier.run(sch.factors, sch.checks, sch.order, 0, syn, 2 * d * d, priority=frontier.priority_of(sch))
    lp_e, cfg_e = emulator.run(sch, syn)
    lp_c, cfg_c = cref.FrontierPlan(sch).run(syn)
    assert np.array_equal(lp_f, lp_e) and np.array_equal(cfg_f, cfg_e)
    assert np.array_equal(lp_f, lp_c) and np.array_equal(cfg_f, cfg_c)
    nq, s2q, pix, pri = networks.general_problem_css(t, em.px, em.py, em.pz)
    lp_d, cfg_d = cref.DensePlan(networks.tnmap_network(nq, s2q, pix, pri), len(s2q), nq, True).run(syn)
    assert np.allclose(lp_f, lp_d, rtol=1e-12, atol=0)
    assert np.array_equal(cfg_f, cfg_d)                         # random per-qubit noise: the maximiser is unique


def test_orders_do_not_change_the_value(tq):
    t = tq.CSSTannerGraph(tq.SurfaceCode(3, 3))
    rng = np.random.default_rng(0)
    em = tq.IndependentDepolarizingError(rng.uniform(0.01, 0.1, 9), rng.uniform(0.01, 0.1, 9), rng.uniform(0.01, 0.1, 9))
    gdp, _ = tq.reduce2general(t, em)
    syn = ((np.arange(256)[:, None] >> np.arange(8)) & 1).astype(np.uint8)
    base = None
    for order in (None, tq.NoOptimizer(), list(rng.permutation(9)), list(range(8, -1, -1))):
        sch = tq.tnmap_schedule(tq.TNMAP(optimizer=order), gdp)
        lp, cfg = emulator.run(sch, syn)
        if base is None:
            base = (lp, cfg)
        assert np.allclose(lp, base[0], rtol=1e-13) and np.array_equal(cfg, base[1])


def test_correlated_and_overlapping_priors(tq):
    rng = np.random.default_rng(5)
    t = tq.CSSTannerGraph(tq.SteaneCode())
    n = 7
    # one rank-4 prior on (x0, z0, x1, z1), overlapping rank-2 priors on a chain, a variable no prior mentions
    ixs = [[0, 7, 1, 8], [2, 9], [9, 3], [3, 10], [4, 11], [5, 12], [6]]
    tensors = [rng.uniform(0.01, 1.0, size=(2,) * len(ix)) for ix in ixs]
    tn = tq.SimpleTensorNetwork(ixs, tensors)
    gdp, _ = tq.reduce2general(t, tn)
    sch = tq.tnmap_schedule(tq.TNMAP(), gdp)
    syn = ((np.arange(64)[:, None] >> np.arange(6)) & 1).astype(np.uint8)
    lp, cfg = emulator.run(sch, syn)
    lp2, cfg2 = frontier.run(sch.factors, sch.checks, sch.order, 0, syn, 14, priority=frontier.priority_of(sch))
    assert np.array_equal(lp, lp2) and np.array_equal(cfg, cfg2)
    en = bruteforce.Enumeration(14, [list(c) for c in gdp.tanner.s2q], ixs, tensors)
    for b in range(64):
        best, maxs = en.map(syn[b])
        assert abs(lp[b] - best) <= 1e-12 * abs(best)
        assert any(np.array_equal(cfg[b], m) for m in maxs)


def test_infeasible_and_orphan_checks(tq):
    # a check with no bits: syndrome bit 1 on it is infeasible (the dense parity tensor gives weight 0)
    tg = tq.SimpleTannerGraph(3, [[0, 1], [], [1, 2]])
    gdp = tq.GeneralDecodingProblem(tg, tq.SimpleTensorNetwork([[0], [1], [2]], [np.array([0.9, 0.1])] * 3))
    sch = tq.tnmap_schedule(tq.TNMAP(), gdp)
    lp, cfg = emulator.run(sch, np.array([[1, 0, 0], [1, 1, 0], [0, 0, 1]], dtype=np.uint8))
    assert np.isfinite(lp[0]) and cfg[0].tolist() == [1, 0, 0]
    assert lp[1] == -np.inf
    assert np.isfinite(lp[2]) and cfg[2].tolist() == [0, 0, 1]
    # zero-probability priors: log(0) = -inf must propagate without NaN
    gdp0 = tq.GeneralDecodingProblem(tg, tq.SimpleTensorNetwork([[0], [1], [2]], [np.array([1.0, 0.0])] * 3))
    lp0, _ = emulator.run(tq.tnmap_schedule(tq.TNMAP(), gdp0), np.array([[0, 0, 0], [1, 0, 0]], dtype=np.uint8))
    assert lp0[0] == 0.0 and lp0[1] == -np.inf


@pytest.mark.parametrize("code", ["steane", "color488_5", "surface_5"])
def test_tnmmap_css_lowering_vs_dense(tq, code):
    c = {"steane": tq.SteaneCode(), "color488_5": tq.Color488(5), "surface_5": tq.SurfaceCode(5, 5)}[code]
    t = tq.CSSTannerGraph(c)
    n = t.stgx.nq
    rng = np.random.default_rng(1)
    em = tq.IndependentDepolarizingError(rng.uniform(0.01, 0.1, n), rng.uniform(0.01, 0.1, n), rng.uniform(0.01, 0.1, n))
    lx, lz, sch, R, L, FIX = tq.tnmmap_css_schedule(tq.TNMMAP(), tq.IndependentDepolarizingDecodingProblem(t, em))
    syn = _random_syndromes(t, em, 3, 24)
    got = emulator.run(sch, syn)
    got2 = frontier.run(sch.factors, sch.checks, sch.order, 1, syn, 2 * n)
    ref = cref.DensePlan(networks.tnmmap_css_network(t, lx, lz, em.px, em.py, em.pz), syn.shape[1], 2 * n, False).run(syn)
    assert np.allclose(got, ref, rtol=1e-10, atol=0) and np.allclose(got2, ref, rtol=1e-10, atol=0)
    # coset representative: H (R s) = s, and the repair rows move exactly one sector bit each
    Hg = np.zeros((syn.shape[1], 2 * n), dtype=np.uint8)
    Hg[: t.stgx.ns, n:], Hg[t.stgx.ns:, :n] = t.stgx.H, t.stgz.H
    e0 = (syn @ R.T) & 1
    assert np.array_equal((e0 @ Hg.T) & 1, syn)
    assert not ((FIX @ Hg.T) & 1).any()
    assert np.array_equal((L @ FIX.T) & 1, np.eye(L.shape[0], dtype=int))


def test_dem_lowering_vs_dense_and_bruteforce(tq):
    dem = tq.parse_dem_file(os.path.join(ROOT, "tests", "golden", "dem.dem"))
    tanner, l2q, sch, R, L, FIX = tq.tnmmap_dem_schedule(tq.TNMMAP(), dem)
    e = philox.sample_flips(dem.error_rates, 1, 0, 40)
    syn = gf2.syndrome_extraction(e, tanner.H)
    got = emulator.run(sch, syn)
    en = bruteforce.Enumeration(21, tanner.s2q, [[i] for i in range(21)], [np.array([1 - p, p]) for p in dem.error_rates])
    for b in range(40):
        assert np.allclose(got[b], en.marginal(syn[b], L), rtol=1e-10, atol=0)
    for fac in (True, False):
        net = networks.tnmmap_dem_network(dem.error_rates, dem.flipped_detectors, 6, 1, factorize=fac)
        assert np.allclose(got, cref.DensePlan(net, 6, 21, False).run(syn), rtol=1e-10, atol=0)
    assert np.array_equal(((syn @ R.T & 1) @ tanner.H.T) & 1, syn)
    assert not ((FIX @ tanner.H.T) & 1).any() and ((L @ FIX.T) & 1).tolist() == [[1]]


def test_schedule_cost_table(tq):
    """Frontier width / candidate evaluations per shot of the BASELINE configs (DESIGN.md quotes these)."""
    expect = {3: (3, 104), 5: (5, 1208), 7: (8, 9976), 9: (10, 68600)}
    for d, (w, cost) in expect.items():
        t = tq.CSSTannerGraph(tq.SurfaceCode(d, d))
        gdp, _ = tq.reduce2general(t, tq.iid_error(0.05, t))
        sch = tq.tnmap_schedule(tq.TNMAP(), gdp)
        assert sch.w_max <= w and sch.cost <= cost


def test_cabi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "tqec.h")).read()
    declared = set(re.findall(r"\b(tqec_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"tqec_plan_desc", "tqec_mc_desc"}
    from tensorqec.jl_b200 import _cabi
    assert declared == set(_cabi.EXPORTS)
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert lib.tqec_version() == 100


def test_fails_loudly_without_gpu(tq):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    t = tq.CSSTannerGraph(tq.SurfaceCode(3, 3))
    with pytest.raises(tq.TqecError):
        tq.compile(tq.TNMAP(), t)
    with pytest.raises(tq.TqecError):
        tq.syndrome_extraction(np.zeros(9, dtype=np.uint8), t.stgz)
    with pytest.raises(tq.TqecError):
        tq.random_error_pattern(tq.iid_error(0.1, 10))


def test_sector_fixes_span_every_reachable_flip():
    """gf2_sector_fixes: undetectable patterns whose sector flips are a reduced-echelon basis of ALL reachable flips,
    joint flips of several observables included (the per-observable solve it replaces missed those)."""
    from tensorqec.jl_b200.tanner import gf2_sector_fixes
    rng = np.random.RandomState(0)
    joint_only = 0
    for _ in range(120):
        n, m, k = rng.randint(4, 12), rng.randint(1, 6), rng.randint(1, 4)
        H = rng.randint(0, 2, (m, n)).astype(np.uint8)
        L = rng.randint(0, 2, (k, n)).astype(np.uint8)
        F = gf2_sector_fixes(H, L)
        assert not ((H.astype(int) @ F.T.astype(int)) & 1).any()
        D = (F.astype(int) @ L.T.astype(int)) & 1
        piv = [int(np.flatnonzero(d)[0]) for d in D]
        assert piv == sorted(set(piv)) and all(D[:, pv].sum() == 1 for pv in piv)
        X = ((np.arange(1 << n)[:, None] >> np.arange(n)) & 1).astype(int)
        ker = X[~((X @ H.T.astype(int)) & 1).any(axis=1)]
        reach = {tuple(r) for r in (ker @ L.T.astype(int)) & 1}
        assert len(reach) == 1 << len(F)
        unit = [tuple(int(i == l) for i in range(k)) for l in range(k)]
        if any(sum(r) > 1 for r in reach) and not all(u in reach for u in unit if any(r[unit.index(u)] for r in reach)):
            joint_only += 1
    assert joint_only > 0, "the fuzz never produced a joint-only flip"


def test_table_decoder_enumeration_matches_the_reference_loop(tq, tmp_path):
    """make_table (truthtable.jl:51-90) vectorised vs the reference's triple loop written out: weights 0..d in order,
    qubit subsets lexicographic, per-qubit Paulis X, Y, Z; a syndrome keeps its first pattern unless a later one is
    strictly more probable.  save_table / load_table round trip (truthtable.jl:138-166)."""
    import itertools
    t = tq.CSSTannerGraph(tq.SurfaceCode(3, 3))
    rng = np.random.default_rng(1)
    em = tq.IndependentDepolarizingError(rng.uniform(0.01, 0.08, 9), rng.uniform(0.01, 0.08, 9), rng.uniform(0.01, 0.08, 9))
    for pvec in (tq.iid_error(0.05, t), em, None):
        tb = tq.make_table(t, 2, pvec)
        Hx, Hz = t.stgx.H.astype(int), t.stgz.H.astype(int)

        def prob(x, z):
            if pvec is None:
                return 0.0
            p = 1.0
            for i in range(9):
                p = p * [[1 - pvec.px[i] - pvec.py[i] - pvec.pz[i], pvec.pz[i]], [pvec.px[i], pvec.py[i]]][x[i]][z[i]]
            return p
        best = {}
        for k in range(3):
            for combo in itertools.combinations(range(9), k):
                for i in range(3 ** k):
                    x, z = np.zeros(9, dtype=int), np.zeros(9, dtype=int)
                    for j, q in enumerate(combo):
                        dg = (i // 3 ** j) % 3
                        x[q], z[q] = dg <= 1, dg >= 1
                    key = tuple(np.concatenate([(Hx @ z) % 2, (Hz @ x) % 2]))
                    if key not in best or prob(*best[key]) < prob(x, z):
                        best[key] = (x, z)
        assert len(tb) == len(best)
        for kb, ev in zip(tq.unpack_bits(tb.keys, 8), tq.unpack_bits(tb.values, 18)):
            x, z = best[tuple(kb)]
            assert np.array_equal(ev[:9], x) and np.array_equal(ev[9:], z)
    f = str(tmp_path / "table.txt")
    tq.save_table(tb, f)
    tb2 = tq.load_table(f, 9, 8)
    assert np.array_equal(tb2.keys, tb.keys) and np.array_equal(tb2.values, tb.values)
