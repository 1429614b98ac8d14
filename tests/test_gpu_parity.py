"""Parity of the CUDA path (through the C ABI) against the oracle.  Integer / index results are compared bit for bit;
FP64 marginals within 1e-10 relative (the tolerance BASELINE.json's north_star states); MAP log-weights bit for bit
(each candidate is one IEEE add and max is exact, so the recurrence is reproducible)."""
import numpy as np
import pytest

from oracle import bruteforce, cref, emulator, frontier, gf2, networks, philox

pytestmark = pytest.mark.gpu

MAR_RTOL = 1e-10


def _css_case(tq, code, p=0.05, pvec=None):
    t = tq.CSSTannerGraph(code)
    em = tq.iid_error(p, t) if pvec is None else pvec
    return t, em


def _syndromes(t, em, seed, B):
    ex, ez = philox.sample_depolarizing(em.px, em.py, em.pz, seed, 0, B)
    sx, sz = gf2.css_syndrome(ex, ez, t.stgx.H, t.stgz.H)
    return ex, ez, sx, sz


def test_tnmap_d3_all_syndromes_exhaustive(tq):
    t, em = _css_case(tq, tq.SurfaceCode(3, 3))
    ct = tq.compile(tq.TNMAP(), t, em)
    syn = ((np.arange(256)[:, None] >> np.arange(8)) & 1).astype(np.uint8)
    res = tq.decode(ct, tq.CSSSyndrome(syn[:, :4], syn[:, 4:]))
    sch = ct.cd.schedule
    lp, cfg = frontier.run(sch.factors, sch.checks, sch.order, 0, syn, 18, priority=frontier.priority_of(sch))
    got = np.concatenate([res.error_pattern.xerror, res.error_pattern.zerror], axis=1)
    assert np.array_equal(got, cfg)
    assert np.array_equal(res.logp, lp)
    # against exhaustive enumeration: the value is the maximum, the pattern is one of the maximisers
    nq, s2q, pix, pri = networks.general_problem_css(t, em.px, em.py, em.pz)
    en = bruteforce.Enumeration(nq, s2q, pix, pri)
    for b in range(256):
        best, maxs = en.map(syn[b])
        assert abs(best - lp[b]) <= 1e-12 * abs(best)
        assert any(np.array_equal(got[b], m) for m in maxs)


@pytest.mark.parametrize("d,B", [(5, 4096), (7, 2048), (9, 1024)])
def test_tnmap_surface_matches_oracle_bit_exact(tq, d, B):
    t, em = _css_case(tq, tq.SurfaceCode(d, d))
    ct = tq.compile(tq.TNMAP(), t, em)
    ex, ez, sx, sz = _syndromes(t, em, 1000 * d, B)
    res = tq.decode(ct, tq.CSSSyndrome(sx, sz))
    sch = ct.cd.schedule
    lp, cfg = cref.FrontierPlan(sch).run(np.concatenate([sx, sz], axis=1))
    n = d * d
    assert np.array_equal(res.error_pattern.xerror, cfg[:, :n])
    assert np.array_equal(res.error_pattern.zerror, cfg[:, n:])
    assert np.array_equal(res.logp, lp)
    # the numpy statement of the recurrence agrees with its C port on a subset
    lp2, cfg2 = frontier.run(sch.factors, sch.checks, sch.order, 0, np.concatenate([sx, sz], axis=1)[:128], 2 * n, priority=frontier.priority_of(sch))
    assert np.array_equal(lp2, lp[:128]) and np.array_equal(cfg2, cfg[:128])
    # decoded pattern reproduces the syndrome (the reference's own assertion, test/decoding/tndecoder.jl:57)
    assert tq.syndrome_extraction(res.error_pattern, t) == tq.CSSSyndrome(sx, sz)


@pytest.mark.parametrize("d", [7, 9])
def test_sweep_kernel_edge_cases(tq, d, monkeypatch):
    """k_sweep (in-place patch sweep): per-qubit noise (unique maximisers), batch sizes that do not fill a pass / a group
    of 32 shots, the all-zero and all-one syndromes, and bit-identity with the general kernels on the same unfused
    schedule (TQEC_NO_SWEEP routes the plan through k_frontier_warp)."""
    from tensorqec.jl_b200 import _cabi
    rng = np.random.default_rng(d)
    n = d * d
    em = tq.IndependentDepolarizingError(rng.uniform(0.01, 0.1, n), rng.uniform(0.01, 0.1, n), rng.uniform(0.01, 0.1, n))
    t = tq.CSSTannerGraph(tq.SurfaceCode(d, d))
    ct = tq.compile(tq.TNMAP(), t, em)
    sch = ct.cd.schedule
    assert ct.cd.plan.query(_cabi.Q_SWEEP) == 1 and getattr(sch, "sweep", None) is not None
    ex, ez, sx, sz = _syndromes(t, em, 77, 1000)
    syn = np.concatenate([sx, sz], axis=1)
    syn[0] = 0
    syn[1] = 1
    oracle_plan = cref.FrontierPlan(sch)
    for B in (1, 2, 31, 33, 1000):
        corr, logp = ct.cd.plan.decode_map(tq.pack_bits(syn[:B]))
        lp, cfg = oracle_plan.run(syn[:B])
        assert np.array_equal(tq.unpack_bits(corr, 2 * n), cfg) and np.array_equal(logp, lp)
    corr, logp = ct.cd.plan.decode_map(tq.pack_bits(syn))
    # same schedule through the general kernels
    monkeypatch.setenv("TQEC_NO_SWEEP", "1")
    plan2 = _cabi.Plan(sch, 0)
    assert plan2.query(_cabi.Q_SWEEP) == 0
    corr2, logp2 = plan2.decode_map(tq.pack_bits(syn))
    assert np.array_equal(corr, corr2) and np.array_equal(logp, logp2)
    plan2.close()


@pytest.mark.parametrize("dx,dz", [(6, 6), (8, 8), (5, 7)])
def test_sweep_kernel_even_and_rectangular_codes(tq, dx, dz, monkeypatch):
    """k_sweep on even-distance and rectangular rotated surface codes: their sweeps open checks into dead slots (fresh
    pins: a pinned variable whose flip mask contains its own bit) and close checks without a successor.  Bit-identical to
    the C port of the recurrence and to the general kernels on the same unfused schedule; default (p, p, p) noise and
    per-qubit noise."""
    from tensorqec.jl_b200 import _cabi
    n = dx * dz
    t = tq.CSSTannerGraph(tq.SurfaceCode(dx, dz))
    rng = np.random.default_rng(dx * 10 + dz)
    for em in (tq.iid_error(0.05, t),
               tq.IndependentDepolarizingError(rng.uniform(0.01, 0.1, n), rng.uniform(0.01, 0.1, n), rng.uniform(0.01, 0.1, n))):
        ct = tq.compile(tq.TNMAP(), t, em)
        sch = ct.cd.schedule
        assert ct.cd.plan.query(_cabi.Q_SWEEP) == 1 and getattr(sch, "sweep", None) is not None
        ex, ez, sx, sz = _syndromes(t, em, 78, 600)
        syn = np.concatenate([sx, sz], axis=1)
        syn[0] = 0
        syn[1] = 1
        oracle_plan = cref.FrontierPlan(sch)
        for B in (1, 33, 600):
            corr, logp = ct.cd.plan.decode_map(tq.pack_bits(syn[:B]))
            lp, cfg = oracle_plan.run(syn[:B])
            assert np.array_equal(tq.unpack_bits(corr, 2 * n), cfg) and np.array_equal(logp, lp)
        res = tq.decode(ct, tq.CSSSyndrome(sx, sz))
        assert tq.syndrome_extraction(res.error_pattern, t) == tq.CSSSyndrome(sx, sz)
        corr, logp = ct.cd.plan.decode_map(tq.pack_bits(syn))
        monkeypatch.setenv("TQEC_NO_SWEEP", "1")
        plan2 = _cabi.Plan(sch, 0)
        monkeypatch.delenv("TQEC_NO_SWEEP")
        assert plan2.query(_cabi.Q_SWEEP) == 0
        corr2, logp2 = plan2.decode_map(tq.pack_bits(syn))
        assert np.array_equal(corr, corr2) and np.array_equal(logp, logp2)
        plan2.close()


@pytest.mark.parametrize("extra", [5_000, 30_000])
def test_sweep_tail_launch_does_not_change_results(tq, extra, monkeypatch):
    """A batch that ends inside a round is decoded by two launches of k_sweep -- whole rounds with 32 shots per team and
    round, the tail with 8 or 16 -- and gives the same corrections and log-weights, bit for bit, as a single launch."""
    from tensorqec.jl_b200 import _cabi
    t, em = _css_case(tq, tq.SurfaceCode(7, 7))
    ct = tq.compile(tq.TNMAP(), t, em)
    assert ct.cd.plan.query(_cabi.Q_SWEEP) == 1
    monkeypatch.setenv("TQEC_PIPE_CHUNK_LOG2", "24")           # one chunk: the host pipeline would cut the batch at whole rounds
    per_round = 148 * 16 * 32
    B = 4 * per_round + extra
    ex, ez, sx, sz = _syndromes(t, em, 31, 8192)
    bits = np.tile(np.concatenate([sx, sz], axis=1), ((B + 8191) // 8192, 1))[:B]
    words = tq.pack_bits(bits)
    l0 = ct.cd.plan.query(_cabi.Q_LAUNCHES)
    corr, logp = ct.cd.plan.decode_map(words)
    launches = ct.cd.plan.query(_cabi.Q_LAUNCHES) - l0
    monkeypatch.setenv("TQEC_SWEEP_NO_TAIL_SPLIT", "1")
    l1 = ct.cd.plan.query(_cabi.Q_LAUNCHES)
    corr1, logp1 = ct.cd.plan.decode_map(words)
    assert ct.cd.plan.query(_cabi.Q_LAUNCHES) - l1 == 1
    assert launches == 2, "the tail should have been launched separately (148 SMs x 16 teams x 32 shots per round)"
    assert np.array_equal(corr, corr1) and np.array_equal(logp, logp1)
    lp, cfg = cref.FrontierPlan(ct.cd.schedule).run(bits[-2000:])
    assert np.array_equal(tq.unpack_bits(corr[-2000:], 98), cfg) and np.array_equal(logp[-2000:], lp)


def test_sweep_kernel_long_head(tq):
    """The 12-bit tabulated head bench.py uses (17 of the 81 steps of d = 9 become a table look-up): same corrections and
    log-weights, bit for bit, as the C port of the full recurrence and as the default 10-bit head."""
    t, em = _css_case(tq, tq.SurfaceCode(9, 9))
    ex, ez, sx, sz = _syndromes(t, em, 4242, 700)
    syn = tq.CSSSyndrome(sx, sz)
    res = {}
    for bits in (10, 12):
        ct = tq.compile(tq.TNMAP(head_bits=bits), t, em)
        assert len(ct.cd.schedule.sweep.head_bits) == bits
        res[bits] = tq.decode(ct, syn)
    lp, cfg = cref.FrontierPlan(ct.cd.schedule).run(np.concatenate([sx, sz], axis=1))
    for bits in (10, 12):
        got = np.concatenate([res[bits].error_pattern.xerror, res[bits].error_pattern.zerror], axis=1)
        assert np.array_equal(got, cfg) and np.array_equal(res[bits].logp, lp)


def test_tnmap_matches_dense_reference_value(tq):
    """MAP value equals the dense (reference-style) contraction; the pattern has that weight and the syndrome."""
    d = 5
    rng = np.random.default_rng(d)
    n = d * d
    em = tq.IndependentDepolarizingError(rng.uniform(0.01, 0.1, n), rng.uniform(0.01, 0.1, n), rng.uniform(0.01, 0.1, n))
    t = tq.CSSTannerGraph(tq.SurfaceCode(d, d))
    ct = tq.compile(tq.TNMAP(), t, em)
    ex, ez, sx, sz = _syndromes(t, em, 11, 256)
    res = tq.decode(ct, tq.CSSSyndrome(sx, sz))
    nq, s2q, pix, pri = networks.general_problem_css(t, em.px, em.py, em.pz)
    dp = cref.DensePlan(networks.tnmap_network(nq, s2q, pix, pri), len(s2q), nq, True)
    lp, cfg = dp.run(np.concatenate([sx, sz], axis=1))
    assert np.allclose(res.logp, lp, rtol=1e-12, atol=0)
    # random per-qubit noise: the maximiser is unique, so the patterns coincide with the reference-style contraction
    got = np.concatenate([res.error_pattern.xerror, res.error_pattern.zerror], axis=1)
    assert np.array_equal(got, cfg)


def test_tnmap_d9_matches_dense_reference(tq):
    """The headline size against the reference-style DENSE contraction (oracle C port of the pairwise contraction of the
    unfactorised network on a greedy tree, sc = 16): MAP values agree to rounding; under per-qubit noise (unique
    maximisers) the corrections are identical; under the reference's default (p, p, p) noise, where exact ties are
    common, both corrections have the same weight and reproduce the syndrome, and the fraction of shots on
    which the two tie-breaks land in different logical classes is bounded (DESIGN.md section 2 quotes the measured
    rate: it is a property of the tie rule, not an error of either contraction)."""
    d, n, B = 9, 81, 64
    t = tq.CSSTannerGraph(tq.SurfaceCode(d, d))
    lx, lz = tq.logical_operator(t)
    rng = np.random.default_rng(99)
    generic = tq.IndependentDepolarizingError(rng.uniform(0.01, 0.1, n), rng.uniform(0.01, 0.1, n), rng.uniform(0.01, 0.1, n))
    for em, unique in ((generic, True), (tq.iid_error(0.05, t), False)):
        ct = tq.compile(tq.TNMAP(), t, em)
        ex, ez, sx, sz = _syndromes(t, em, 9, B)
        syn = np.concatenate([sx, sz], axis=1)
        res = tq.decode(ct, tq.CSSSyndrome(sx, sz))
        nq, s2q, pix, pri = networks.general_problem_css(t, em.px, em.py, em.pz)
        dp = cref.DensePlan(networks.tnmap_network(nq, s2q, pix, pri), len(s2q), nq, True)
        lp, cfg = dp.run(syn)
        assert np.allclose(res.logp, lp, rtol=1e-12, atol=0)
        got = np.concatenate([res.error_pattern.xerror, res.error_pattern.zerror], axis=1)
        if unique:
            assert np.array_equal(got, cfg)
            continue
        # ties: same weight (the value of either pattern under the prior equals the MAP value) and same syndrome
        T = np.log(np.array([[1 - 0.15, 0.05], [0.05, 0.05]]))
        for pat in (got, cfg):
            assert np.allclose(T[pat[:, :n], pat[:, n:]].sum(axis=1), lp, rtol=1e-12, atol=0)
            assert np.array_equal(gf2.css_syndrome(pat[:, :n], pat[:, n:], t.stgx.H, t.stgz.H)[0], sx)
            assert np.array_equal(gf2.css_syndrome(pat[:, :n], pat[:, n:], t.stgx.H, t.stgz.H)[1], sz)
        diff = gf2.check_logical_error_css(got[:, :n], got[:, n:], cfg[:, :n], cfg[:, n:], lx, lz)
        same = int((got == cfg).all(axis=1).sum())
        print(f"d=9 (p,p,p): identical corrections {same}/{B}, different logical class {int(diff.sum())}/{B}")
        assert diff.mean() < 0.25


@pytest.mark.parametrize("code", ["steane", "color488_5"])
def test_tnmap_small_codes(tq, code):
    c = tq.SteaneCode() if code == "steane" else tq.Color488(5)
    t, em = _css_case(tq, c)
    ct = tq.compile(tq.TNMAP(), t, em)
    ex, ez, sx, sz = _syndromes(t, em, 5, 1000)
    res = tq.decode(ct, tq.CSSSyndrome(sx, sz))
    sch = ct.cd.schedule
    lp, cfg = cref.FrontierPlan(sch).run(np.concatenate([sx, sz], axis=1))
    n = t.stgx.nq
    assert np.array_equal(np.concatenate([res.error_pattern.xerror, res.error_pattern.zerror], axis=1), cfg)
    assert np.array_equal(res.logp, lp)
    assert tq.syndrome_extraction(res.error_pattern, t) == tq.CSSSyndrome(sx, sz)


@pytest.mark.parametrize("code", ["surface3", "steane", "color488_5"])
def test_tabulated_plans_equal_the_kernels(tq, code, monkeypatch):
    """Plans with <= 16 syndrome bits are decoded once per syndrome at plan creation and served from that table
    (TQEC_Q_TABLE): the table holds the kernels' own outputs, so both paths agree bit for bit -- on every syndrome of the
    code for the small ones -- and with the oracle."""
    from tensorqec.jl_b200 import _cabi
    c = {"surface3": tq.SurfaceCode(3, 3), "steane": tq.SteaneCode(), "color488_5": tq.Color488(5)}[code]
    t, em = _css_case(tq, c)
    ct = tq.compile(tq.TNMAP(), t, em)
    sch = ct.cd.schedule
    assert ct.cd.plan.query(_cabi.Q_TABLE) == 1
    ns = sch.n_checks
    if ns <= 8:
        syn = ((np.arange(1 << ns)[:, None] >> np.arange(ns)) & 1).astype(np.uint8)
    else:
        ex, ez, sx, sz = _syndromes(t, em, 21, 3000)
        syn = np.concatenate([sx, sz], axis=1)
    corr, logp = ct.cd.plan.decode_map(tq.pack_bits(syn))
    monkeypatch.setenv("TQEC_NO_TABLE", "1")
    plan2 = _cabi.Plan(sch, 0)
    assert plan2.query(_cabi.Q_TABLE) == 0
    corr2, logp2 = plan2.decode_map(tq.pack_bits(syn))
    plan2.close()
    assert np.array_equal(corr, corr2) and np.array_equal(logp, logp2)
    lp, cfg = cref.FrontierPlan(sch).run(syn)
    assert np.array_equal(tq.unpack_bits(corr, sch.n_vars), cfg) and np.array_equal(logp, lp)
    # sum-product plans are tabulated the same way
    if code == "surface3":
        monkeypatch.delenv("TQEC_NO_TABLE")
        ctm = tq.compile(tq.TNMMAP(), t, em)
        assert ctm.plan.query(_cabi.Q_TABLE) == 1
        mar, arg = ctm.plan.decode_marginal(tq.pack_bits(syn))
        monkeypatch.setenv("TQEC_NO_TABLE", "1")
        plan3 = _cabi.Plan(ctm.schedule, 0)
        assert plan3.query(_cabi.Q_TABLE) == 0
        mar3, arg3 = plan3.decode_marginal(tq.pack_bits(syn))
        plan3.close()
        assert np.array_equal(mar, mar3) and np.array_equal(arg, arg3)


def test_tnmap_single_shot_and_classical(tq):
    t = tq.CSSTannerGraph(tq.SurfaceCode(3, 3))
    ct = tq.compile(tq.TNMAP(), t.stgz)                      # classical problem, iid_error(0.05, n) default
    e = np.array([0, 0, 0, 0, 1, 0, 0, 0, 0], dtype=np.uint8)
    syn = tq.syndrome_extraction(e, t.stgz)
    res = tq.decode(ct, syn)
    assert res.error_pattern.shape == (9,)
    assert syn == tq.syndrome_extraction(res.error_pattern, t.stgz)
    res2 = tq.decode(tq.TNMAP(), t.stgz, syn)
    assert np.array_equal(res2.error_pattern, res.error_pattern)


def test_tnmap_correlated_prior(tq):
    """test/decoding/general_decoding.jl:24-46: a rank-4 prior tensor on (x4, z4, x7, z7) (0-based qubits 3 and 6)."""
    t = tq.CSSTannerGraph(tq.SurfaceCode(3, 3))
    n = 9
    singles = [0, 1, 2, 4, 5, 7, 8]
    a = np.zeros((2, 2, 2, 2))
    a[0, 0, 0, 0] = 0.99
    a[0, 0, 1, 1] = 0.01
    tn = tq.SimpleTensorNetwork([[i, i + n] for i in singles] + [[3, 3 + n, 6, 6 + n]],
                                [tq.single_qubit_tensor(0.05, 0.03, 0.01) for _ in singles] + [a])
    gdp, red = tq.reduce2general(t, tn)
    ct = tq.compile(tq.TNMAP(), gdp)
    syn = ((np.arange(256)[:, None] >> np.arange(8)) & 1).astype(np.uint8)
    res = tq.decode(ct, tq.SimpleSyndrome(syn))
    nq, s2q = 18, [list(c) for c in gdp.tanner.s2q]
    en = bruteforce.Enumeration(nq, s2q, tn.ixs, tn.tensors)
    for b in range(256):
        best, maxs = en.map(syn[b])
        if np.isfinite(best):
            assert res.success_tag[b]
            assert abs(best - res.logp[b]) <= 1e-12 * abs(best)
            assert any(np.array_equal(res.error_pattern[b], m) for m in maxs)
        else:
            assert not res.success_tag[b]


def test_tnmmap_golden_marginal(tq):
    """test/decoding/tndecoder.jl:101-110: the only numeric contraction golden of the reference."""
    t = tq.CSSTannerGraph(tq.SurfaceCode(3, 3))
    p = np.full(9, 0.1)
    ct = tq.compile(tq.TNMMAP(), t, tq.IndependentDepolarizingError(p, np.zeros(9), np.zeros(9)))
    mar = ct.marginal(tq.CSSSyndrome(np.zeros(4, dtype=np.uint8), np.zeros(4, dtype=np.uint8)))
    assert np.allclose(mar, [[0.3972875040000002, 0.004284496000000001], [0.0, 0.0]], atol=1e-10)


def test_tnmmap_d3_vs_bruteforce_and_error_pattern(tq):
    rng = np.random.default_rng(3)
    t = tq.CSSTannerGraph(tq.SurfaceCode(3, 3))
    em = tq.IndependentDepolarizingError(rng.uniform(0.01, 0.1, 9), rng.uniform(0.01, 0.1, 9), rng.uniform(0.01, 0.1, 9))
    ct = tq.compile(tq.TNMMAP(), t, em)
    syn = ((np.arange(256)[:, None] >> np.arange(8)) & 1).astype(np.uint8)
    css = tq.CSSSyndrome(syn[:, :4], syn[:, 4:])
    res = tq.decode(ct, css)
    lx, lz = tq.logical_operator(t)
    nq, s2q, pix, pri = networks.general_problem_css(t, em.px, em.py, em.pz)
    en = bruteforce.Enumeration(nq, s2q, pix, pri)
    Lg = np.zeros((2, 18), dtype=np.uint8)
    Lg[0, 9:] = lx[0]
    Lg[1, :9] = lz[0]
    for b in range(256):
        ref = en.marginal(syn[b], Lg)
        got = res.marginal[b].reshape(-1, order="F")
        assert np.allclose(got, ref, rtol=MAR_RTOL, atol=0)
        assert res.sector[b] == int(np.argmax(ref))
    # error pattern: reproduces the syndrome and lies in the decoded sector (tndecoder.jl:167-174)
    assert tq.syndrome_extraction(res.error_pattern, t) == css
    a = (res.error_pattern.zerror @ lx[0]) & 1
    bb = (res.error_pattern.xerror @ lz[0]) & 1
    assert np.array_equal(a + 2 * bb, res.sector)


@pytest.mark.parametrize("d", [5, 7])
def test_tnmmap_surface_vs_dense_reference(tq, d):
    t, em = _css_case(tq, tq.SurfaceCode(d, d), pvec=tq.iid_error(0.04, 0.02, 0.06, d * d))
    ct = tq.compile(tq.TNMMAP(), t, em)
    B = 256 if d == 5 else 32
    ex, ez, sx, sz = _syndromes(t, em, 77, B)
    mar = ct.marginal(tq.CSSSyndrome(sx, sz)).reshape(B, -1, order="F")
    lx, lz = tq.logical_operator(t)
    dp = cref.DensePlan(networks.tnmmap_css_network(t, lx, lz, em.px, em.py, em.pz), 2 * t.stgx.ns, 2 * d * d, False)
    ref = dp.run(np.concatenate([sx, sz], axis=1))
    assert np.allclose(mar, ref, rtol=MAR_RTOL, atol=1e-300)


def test_tnmmap_d9_through_the_sweep(tq):
    """TNMMAP (CSS) at d = 9: 10-bit state, one shot per team pass, 10-bit head table, sum-product layers of k_sweep --
    against the C port of the recurrence at the north-star tolerance, and against the general kernels."""
    from tensorqec.jl_b200 import _cabi
    t, em = _css_case(tq, tq.SurfaceCode(9, 9))
    ct = tq.compile(tq.TNMMAP(), t, em)
    assert ct.plan.query(_cabi.Q_SWEEP) == 1
    ex, ez, sx, sz = _syndromes(t, em, 99, 300)
    res = tq.decode(ct, tq.CSSSyndrome(sx, sz))
    sch = ct.schedule
    ref = cref.FrontierPlan(sch).run(np.concatenate([sx, sz], axis=1))
    got = res.marginal.reshape(300, -1, order="F")
    assert np.allclose(got, ref, rtol=MAR_RTOL, atol=0)
    assert np.array_equal(res.sector, np.argmax(ref, axis=1))
    assert tq.syndrome_extraction(res.error_pattern, t) == tq.CSSSyndrome(sx, sz)


@pytest.mark.parametrize("dx,dz", [(6, 6), (5, 7), (6, 8)])
def test_tnmmap_even_and_rectangular_codes_through_the_sweep(tq, dx, dz):
    """TNMMAP (CSS) of even-distance and rectangular codes on k_sweep<SUMPROD> (fresh pins, extended instantiation):
    against the C port of the recurrence at the north-star tolerance."""
    from tensorqec.jl_b200 import _cabi
    t, em = _css_case(tq, tq.SurfaceCode(dx, dz))
    ct = tq.compile(tq.TNMMAP(), t, em)
    assert ct.plan.query(_cabi.Q_SWEEP) == 1
    ex, ez, sx, sz = _syndromes(t, em, 100 + dx, 300)
    res = tq.decode(ct, tq.CSSSyndrome(sx, sz))
    sch = ct.schedule
    ref = cref.FrontierPlan(sch).run(np.concatenate([sx, sz], axis=1))
    got = res.marginal.reshape(300, -1, order="F")
    assert np.allclose(got, ref, rtol=MAR_RTOL, atol=0)
    # the decoded sector is the argmax wherever it is unique ((p, p, p) noise makes some sectors tie exactly, and an FMA
    # rounds such a pair differently from the C port)
    srt = np.sort(ref, axis=1)
    clear = srt[:, -1] > srt[:, -2] * (1 + 1e-9)
    assert clear.sum() > 200 and np.array_equal(res.sector[clear], np.argmax(ref, axis=1)[clear])
    assert np.array_equal(res.sector, np.argmax(got, axis=1))
    assert tq.syndrome_extraction(res.error_pattern, t) == tq.CSSSyndrome(sx, sz)


def test_byte_entry_points_multi_chunk_pipeline(tq):
    """tqec_decode_map_bytes / tqec_decode_marginal_bytes on batches that span several chunks of the pinned staging
    pipeline (2^19 shots per chunk, two slots, three streams): same results as the packed-word entry points."""
    B = 1_300_000
    t, em = _css_case(tq, tq.SurfaceCode(5, 5))
    ex, ez, sx, sz = _syndromes(t, em, 5, 4096)
    rep = (B + 4095) // 4096
    bits = np.tile(np.concatenate([sx, sz], axis=1), (rep, 1))[:B].copy()
    bits[-1] ^= 1                                                # the last shot differs from its tile-mates
    ct = tq.compile(tq.TNMAP(), t, em)
    corr_w, logp_w = ct.cd.plan.decode_map(tq.pack_bits(bits))
    corr_b, logp_b = ct.cd.plan.decode_map_bits(bits, 50)
    assert np.array_equal(tq.unpack_bits(corr_w, 50), corr_b) and np.array_equal(logp_w, logp_b)
    corr_2, logp_2 = ct.cd.plan.decode_map_bits2(bits[:, :12].copy(), bits[:, 12:].copy(), 50)   # sx | sz as two arrays
    assert np.array_equal(corr_2, corr_b) and np.array_equal(logp_2, logp_b)
    cm = tq.compile(tq.TNMMAP(), t, em)
    mar_w, arg_w = cm.plan.decode_marginal(tq.pack_bits(bits))
    mar_b, arg_b = cm.plan.decode_marginal_bits(bits)
    assert np.array_equal(mar_w, mar_b) and np.array_equal(arg_w, arg_b)


def test_gf2_kernels_bit_exact(tq):
    rng = np.random.default_rng(0)
    for rows, cols in [(2, 5), (40, 81), (80, 162), (130, 300), (1, 1)]:
        H = rng.integers(0, 2, size=(rows, cols)).astype(np.uint8)
        e = rng.integers(0, 2, size=(777, cols)).astype(np.uint8)
        assert np.array_equal(tq.syndrome_extraction(e, H).s, gf2.syndrome_extraction(e, H))
    # the reference's known answers (test/codes/ldpc.jl:34-39, test/decoding/error_model.jl:21-27)
    tg = tq.SimpleTannerGraph(5, [[0, 1, 2, 3], [1, 2, 3, 4]])
    assert list(tq.syndrome_extraction(np.array([1, 0, 1, 1, 0]), tg).s) == [1, 0]
    t = tq.CSSTannerGraph(tq.SurfaceCode(3, 3))
    lx, lz = tq.logical_operator(t)
    z9 = np.zeros(9, dtype=np.uint8)
    assert tq.check_logical_error(z9, np.array([1, 1, 1, 0, 0, 0, 0, 0, 0]), lz) is True
    assert tq.check_logical_error(z9, np.array([1, 1, 0, 1, 1, 0, 0, 0, 0]), lz) is False
    assert tq.check_logical_error(tq.CSSErrorPattern(z9, z9), tq.CSSErrorPattern(z9, z9), lx, lz) is False


def test_sampler_bit_exact_and_offset_invariant(tq):
    em = tq.iid_error(0.05, 0.06, 0.1, 81)
    ep = tq.random_error_pattern(em, seed=12345, shots=5000)
    ox, oz = philox.sample_depolarizing(em.px, em.py, em.pz, 12345, 0, 5000)
    assert np.array_equal(ep.xerror, ox) and np.array_equal(ep.zerror, oz)
    ep2 = tq.random_error_pattern(em, seed=12345, shots=1000, shot_offset=4000)
    assert np.array_equal(ep2.xerror, ox[4000:]) and np.array_equal(ep2.zerror, oz[4000:])
    # rates (test/decoding/error_model.jl:10-19)
    assert abs(ep.xerror.mean() - 0.11) < 0.01 and abs(ep.zerror.mean() - 0.16) < 0.01
    fl = tq.iid_error(0.1, 100)
    assert np.array_equal(tq.random_error_pattern(fl, seed=(1 << 40) + 5, shots=300), philox.sample_flips(fl.p, (1 << 40) + 5, 0, 300))
    one = tq.random_error_pattern(fl, seed=3)
    assert one.shape == (100,)


@pytest.mark.parametrize("n", [5, 64, 65, 128, 200])
def test_sampler_per_qubit_rates_word_boundaries(tq, n):
    """k_sample_depol (one thread per shot, one draw per qubit, integer thresholds) against the oracle's float compare:
    per-qubit rates, z blocks that straddle word boundaries, p = 0 and px + py + pz = 1 sites."""
    rng = np.random.default_rng(n)
    px, py, pz = rng.uniform(0, 0.3, n), rng.uniform(0, 0.3, n), rng.uniform(0, 0.3, n)
    px[0] = py[0] = pz[0] = 0.0
    px[-1], py[-1], pz[-1] = 0.25, 0.25, 0.5
    em = tq.IndependentDepolarizingError(px, py, pz)
    ep = tq.random_error_pattern(em, seed=(7 << 33) + n, shots=3000, shot_offset=(1 << 32) - 100)
    ox, oz = philox.sample_depolarizing(px, py, pz, (7 << 33) + n, (1 << 32) - 100, 3000)
    assert np.array_equal(ep.xerror, ox) and np.array_equal(ep.zerror, oz)
    assert not ep.xerror[:, 0].any() and not ep.zerror[:, 0].any() and (ep.xerror[:, -1] | ep.zerror[:, -1]).all()


@pytest.mark.parametrize("d,shots", [(3, 20000), (5, 20000), (7, 6000)])
def test_fused_pipeline_counts_bit_exact(tq, d, shots):
    t, em = _css_case(tq, tq.SurfaceCode(d, d))
    mc = tq.MonteCarlo(t, tq.TNMAP(), em)
    counts, ms = mc.run(shots, seed=42, chunk=6000)
    ex, ez, sx, sz = _syndromes(t, em, 42, shots)
    sch = mc.compiled.cd.schedule
    _, cfg = cref.FrontierPlan(sch).run(np.concatenate([sx, sz], axis=1))
    n = d * d
    lx, lz = tq.logical_operator(t)
    fx = gf2.check_logical_error(ex, cfg[:, :n], lz)
    fz = gf2.check_logical_error(ez, cfg[:, n:], lx)
    assert list(counts) == [int(fx.sum()), int(fz.sum()), int((fx | fz).sum()), shots]
    # splitting the run over shot ranges (what the GPU sharding does) gives the same totals
    c1, _ = mc.run(shots // 2, seed=42)
    c2, _ = mc.run(shots - shots // 2, seed=42, shot_offset=shots // 2)
    assert list(c1 + c2) == list(counts)
    rates = tq.multi_round_qec(t, tq.TNMAP(), em, rounds=shots, seed=42)
    assert rates == (counts[0] / shots, counts[1] / shots, counts[2] / shots)


def test_dem_tnmmap(tq):
    import os
    dem = tq.parse_dem_file(os.path.join(os.path.dirname(__file__), "golden", "dem.dem"))
    ct = tq.compile(tq.TNMMAP(), dem)
    ep = tq.random_error_pattern(dem, seed=12323, shots=512)
    syn = tq.syndrome_extraction(ep, ct.tanner)
    res = tq.decode(ct, syn)
    # the reference's assertion (test/decoding/tndecoder.jl:112-120)
    assert syn == tq.syndrome_extraction(res.error_pattern, ct.tanner)
    # marginals against exhaustive enumeration over the 2^21 mechanism patterns
    flipped = dem.flipped_detectors
    nd = dem.n_detectors
    en = bruteforce.Enumeration(21, ct.tanner.s2q, [[e] for e in range(21)], [np.array([1 - p, p]) for p in dem.error_rates])
    L = np.zeros((1, 21), dtype=np.uint8)
    L[0, ct.l2q[0]] = 1
    for b in range(64):
        ref = en.marginal(syn.s[b], L)
        assert np.allclose(res.marginal[b], ref, rtol=MAR_RTOL, atol=0)
        assert res.sector[b] == int(np.argmax(ref))
    # dense reference network with and without the rank-3 factorisation (tndecoder.jl:221-238)
    for fac in (True, False):
        net = networks.tnmmap_dem_network(dem.error_rates, flipped, nd, 1, factorize=fac)
        ref = cref.DensePlan(net, nd, 21, False).run(syn.s[:64])
        assert np.allclose(res.marginal[:64].reshape(64, -1), ref, rtol=MAR_RTOL, atol=0)


@pytest.mark.parametrize("name,n_dense", [("surface_d3_r3_phenom.dem", 64), ("surface_d5_r5_phenom.dem", 0)])
def test_dem_surface_memory(tq, name, n_dense):
    """BASELINE configs[3] shape: surface-code memory DEMs (phenomenological, generated by benchmarks/make_dem.py because
    stim is not available).  d=5 x 5 rounds has an 11-bit frontier: since round 2 a plan of rank-1 factors that wide runs
    on the butterfly executor (k_wide_bf, one tile per shot)."""
    import os
    dem = tq.parse_dem_file(os.path.join(os.path.dirname(__file__), "golden", name))
    ct = tq.compile(tq.TNMMAP(), dem)
    B = 512
    ep = tq.random_error_pattern(dem, seed=4, shots=B)
    syn = tq.syndrome_extraction(ep, ct.tanner)
    res = tq.decode(ct, syn)
    assert syn == tq.syndrome_extraction(res.error_pattern, ct.tanner)
    sch = ct.schedule
    from tensorqec.jl_b200 import _cabi, schedule as S
    if name.startswith("surface_d5"):
        assert ct.plan.query(_cabi.Q_WIDE) == 1, "an 11-bit plan of rank-1 factors should run as register butterflies (k_wide_bf)"
    # the C port executes the on-chip step tables: lower them for the same order when the plan itself went to the
    # global-memory executor
    csch = sch if hasattr(sch, "hdr") else S.lower(sch.factors, sch.checks, S.SUMPROD, sch.n_vars, sch.n_checks, sch.n_obs, order=sch.order)
    ref = cref.FrontierPlan(csch).run(syn.s)
    got = res.marginal.reshape(B, -1, order="F")
    assert np.allclose(got, ref, rtol=MAR_RTOL, atol=0)
    ref2 = frontier.run(sch.factors, sch.checks, sch.order, 1, syn.s[:8], sch.n_vars)
    assert np.allclose(got[:8], ref2, rtol=MAR_RTOL, atol=0)
    if n_dense:
        net = networks.tnmmap_dem_network(dem.error_rates, dem.flipped_detectors, dem.n_detectors, 1, factorize=True)
        dref = cref.DensePlan(net, dem.n_detectors, len(dem.error_rates), False).run(syn.s[:n_dense])
        assert np.allclose(got[:n_dense], dref, rtol=MAR_RTOL, atol=0)
    # the decoded observable agrees with the true one far more often than not at p = 1 %
    true_obs = (ep[:, ct.l2q[0]].sum(axis=1) & 1)
    assert (res.sector == true_obs).mean() > 0.9


def test_dem_from_generated_circuit(tq):
    """BASELINE configs[3] proper: d=3 x 3 rounds rotated-surface memory circuit with circuit-level noise (depolarizing
    after every Clifford, data depolarizing per round, measurement and reset flips) -> circuit.detector_error_model ->
    TNMMAP on the GPU (12-bit frontier: one tile of the global-memory executor by default, the on-chip CTA kernels when
    asked) against the C port of the recurrence and its numpy statement."""
    txt = tq.surface_memory_circuit(3, 3, "Z", after_clifford_depolarization=2e-3, before_round_data_depolarization=2e-3,
                                    before_measure_flip_probability=2e-3, after_reset_flip_probability=2e-3)
    dem = tq.detector_error_model(tq.parse_stim_string(txt))
    assert dem.n_detectors == 24 and dem.n_observables == 1
    import os
    from tensorqec.jl_b200 import _cabi
    B = 256
    for onchip in (False, True):
        if onchip:
            os.environ["TQEC_SUMPROD_ONCHIP_WIDTH"] = "13"
        try:
            ct = tq.compile(tq.TNMMAP(table_bits=0), dem)
            sch = ct.schedule
        finally:
            os.environ.pop("TQEC_SUMPROD_ONCHIP_WIDTH", None)
        assert ct.plan.query(_cabi.Q_WIDE) == (0 if onchip else 1)
        ep = tq.random_error_pattern(dem, seed=11, shots=B)
        syn = tq.syndrome_extraction(ep, ct.tanner)
        res = tq.decode(ct, syn)
        assert syn == tq.syndrome_extraction(res.error_pattern, ct.tanner)
        got = res.marginal.reshape(B, -1, order="F")
        if onchip:
            ref = cref.FrontierPlan(sch).run(syn.s)
            assert np.allclose(got, ref, rtol=MAR_RTOL, atol=0)
        ref2 = frontier.run(sch.factors, sch.checks, sch.order, 1, syn.s[:8], sch.n_vars)
        assert np.allclose(got[:8], ref2, rtol=MAR_RTOL, atol=0)
        true_obs = (ep[:, ct.l2q[0]].sum(axis=1) & 1)
        assert (res.sector == true_obs).mean() > 0.97


@pytest.mark.parametrize("t_max", [8, 12])
def test_wide_executor_circuit_level_d3(tq, monkeypatch, t_max):
    """The global-memory executor (k_wide_pass) on a plan the on-chip kernels also run: circuit-level d=3 x 3 rounds DEM
    forced through the wide lowering, with 8-bit tiles (passes with up to 4 spectator bits: 16 tiles per shot) and with
    12-bit tiles, against the recurrence oracle, its emulator and the on-chip kernels."""
    from oracle import wide_emulator
    from tensorqec.jl_b200 import _cabi
    txt = tq.surface_memory_circuit(3, 3, "Z", after_clifford_depolarization=2e-3, before_round_data_depolarization=2e-3,
                                    before_measure_flip_probability=2e-3, after_reset_flip_probability=2e-3)
    dem = tq.detector_error_model(tq.parse_stim_string(txt))
    monkeypatch.setenv("TQEC_SUMPROD_ONCHIP_WIDTH", "13")        # this 12-bit plan defaults to the global-memory executor
    on_chip = tq.compile(tq.TNMMAP(table_bits=0), dem)
    assert on_chip.plan.query(_cabi.Q_WIDE) == 0
    monkeypatch.delenv("TQEC_SUMPROD_ONCHIP_WIDTH")
    monkeypatch.setenv("TQEC_FORCE_WIDE", "1")
    monkeypatch.setenv("TQEC_WIDE_TMAX", str(t_max))
    ct = tq.compile(tq.TNMMAP(table_bits=0), dem)
    assert ct.plan.query(_cabi.Q_WIDE) == 1 and ct.plan.query(_cabi.Q_TABLE) == 0
    B = 300
    ep = tq.random_error_pattern(dem, seed=11, shots=B)
    syn = tq.syndrome_extraction(ep, ct.tanner)
    res = tq.decode(ct, syn)
    assert syn == tq.syndrome_extraction(res.error_pattern, ct.tanner)
    wp = ct.schedule
    got = res.marginal.reshape(B, -1, order="F")
    ref = frontier.run(wp.factors, wp.checks, wp.order, 1, syn.s[:16], wp.n_vars)
    assert np.allclose(got[:16], ref, rtol=MAR_RTOL, atol=0)
    emu = wide_emulator.run(wp, syn.s[:16])
    assert np.allclose(got[:16], emu, rtol=1e-13, atol=0)           # same operations in the same order (FMA contraction aside)
    chip = tq.decode(on_chip, syn)
    assert np.allclose(got, chip.marginal.reshape(B, -1, order="F"), rtol=MAR_RTOL, atol=0)
    assert np.array_equal(res.sector, chip.sector)
    # single shot and an empty batch
    one = tq.decode(ct, tq.SimpleSyndrome(syn.s[7]))
    assert np.allclose(one.marginal.reshape(-1, order="F"), got[7], rtol=1e-15, atol=0)
    assert ct.plan.query(_cabi.Q_WIDE_BATCH) >= 1


def test_dem_error_pattern_lands_in_the_reported_sector(tq):
    """Random small DEMs with two observables: the returned pattern reproduces the detectors and, whenever success_tag is
    set, lies in the decoded sector -- also when only a joint flip of both observables is undetectable."""
    rng = np.random.RandomState(4)
    checked = unreachable = 0
    for trial in range(12):
        ne, nd = rng.randint(6, 11), rng.randint(2, 5)
        flips = []
        for e in range(ne):
            det = sorted(set(rng.randint(0, nd, size=rng.randint(1, 3)).tolist()))
            obs = [nd + l for l in range(2) if rng.rand() < 0.4]
            flips.append(det + obs)
        for dd in range(nd):                                      # every detector is touched
            if not any(dd in f for f in flips):
                flips[rng.randint(ne)].append(dd)
        dem = tq.DetectorErrorModel(list(rng.uniform(0.02, 0.3, ne)), [sorted(f) for f in flips], list(range(nd)), [nd, nd + 1])
        ct = tq.compile(tq.TNMMAP(), dem)
        H = ct.tanner.H.astype(int)
        Lm = np.zeros((2, ne), dtype=int)
        for l, c in enumerate(ct.l2q):
            Lm[l, c] = 1
        X = ((np.arange(1 << ne)[:, None] >> np.arange(ne)) & 1).astype(int)
        syn = np.unique((X @ H.T) & 1, axis=0).astype(np.uint8)   # every reachable detector pattern
        res = tq.decode(ct, tq.SimpleSyndrome(syn))
        e = np.asarray(res.error_pattern).astype(int)
        assert np.array_equal((e @ H.T) & 1, syn)
        sec = (e @ Lm.T) & 1
        got = sec[:, 0] | (sec[:, 1] << 1)
        ok = np.asarray(res.success_tag)
        assert np.array_equal(got[ok], np.asarray(res.sector)[ok])
        # success_tag is False only if NO pattern with these detectors lies in the decoded sector (then its weight is 0
        # and it cannot be the argmax of a positive marginal)
        assert ok.all()
        checked += len(syn)
    assert checked > 50


@pytest.mark.parametrize("d", [3, 5, 9])
def test_library_compile_equals_python_lowering(tq, d, monkeypatch):
    """`compile` lowers inside libtqec_cuda.so (tqec_lower -> tqec_plan_from_lowered; also the single call
    tqec_plan_compile a Julia host makes): decodes are bit-identical to a plan uploaded from the Python lowering's
    tables, for TNMAP (max-plus, traceback) and TNMMAP (sum-product), and to the recurrence oracle."""
    import ctypes as C
    from tensorqec.jl_b200 import _cabi, decoding as D, schedule as S
    t, em = _css_case(tq, tq.SurfaceCode(d, d))
    ex, ez, sx, sz = _syndromes(t, em, 77 + d, 500)
    syn = tq.CSSSyndrome(sx, sz)
    lib_map = tq.compile(tq.TNMAP(table_bits=0), t, em)
    lib_mar = tq.compile(tq.TNMMAP(table_bits=0), t, em)
    assert lib_map.cd._schedule is None and lib_mar._schedule is None     # no Python lowering ran
    r_map, r_mar = tq.decode(lib_map, syn), tq.decode(lib_mar, syn)
    monkeypatch.setenv("TQEC_PY_LOWERING", "1")
    py_map = tq.compile(tq.TNMAP(table_bits=0), t, em)
    py_mar = tq.compile(tq.TNMMAP(table_bits=0), t, em)
    monkeypatch.delenv("TQEC_PY_LOWERING")
    assert py_map.cd._schedule is not None
    p_map, p_mar = tq.decode(py_map, syn), tq.decode(py_mar, syn)
    assert np.array_equal(r_map.error_pattern.xerror, p_map.error_pattern.xerror)
    assert np.array_equal(r_map.error_pattern.zerror, p_map.error_pattern.zerror)
    assert np.array_equal(r_map.logp, p_map.logp)
    assert np.array_equal(r_mar.marginal, p_mar.marginal) and np.array_equal(r_mar.sector, p_mar.sector)
    assert lib_map.cd.plan.query(_cabi.Q_SWEEP) == py_map.cd.plan.query(_cabi.Q_SWEEP) == (1 if d >= 5 else 0)
    lp, cfg = cref.FrontierPlan(py_map.cd.schedule).run(np.concatenate([sx, sz], axis=1))
    assert np.array_equal(r_map.logp, lp) and np.array_equal(r_map.error_pattern.xerror, cfg[:, :d * d])
    # the one-call entry point, raw
    gdp, _ = tq.reduce2general(t, em)
    factors, checks = D._tnmap_graph(gdp)
    prob = _cabi.Problem(factors, checks, S.MAXPLUS, gdp.tanner.nq, gdp.tanner.ns, 0, table_bits=-1)
    h = C.c_void_p()
    _cabi.check(_cabi.lib().tqec_plan_compile(C.byref(prob.desc), C.byref(h)))
    words = tq.pack_bits(np.concatenate([sx, sz], axis=1))
    corr = np.zeros((500, max(1, (2 * d * d + 63) // 64)), dtype=np.uint64)
    logp = np.zeros(500)
    _cabi.check(_cabi.lib().tqec_decode_map(h, words.ctypes.data_as(C.c_void_p), 500, corr.ctypes.data_as(C.c_void_p),
                                            logp.ctypes.data_as(C.c_void_p)))
    _cabi.lib().tqec_plan_destroy(h)
    assert np.array_equal(logp, lp) and np.array_equal(tq.unpack_bits(corr, 2 * d * d), cfg)


def test_plan_from_a_saved_lowering_decodes_identically(tq, tmp_path):
    """tqec_lowered_save -> tqec_lowered_load -> tqec_plan_from_lowered: a plan created from the file decodes bit for bit like
    the plan compiled in one go (TNMAP d = 7 on k_sweep, TNMMAP d = 5 marginals)."""
    from tensorqec.jl_b200 import _cabi, decoding as D, schedule as S
    t, em = _css_case(tq, tq.SurfaceCode(7, 7))
    ex, ez, sx, sz = _syndromes(t, em, 5, 700)
    gdp, _ = tq.reduce2general(t, em)
    factors, checks = D._tnmap_graph(gdp)
    prob = _cabi.Problem(factors, checks, S.MAXPLUS, gdp.tanner.nq, gdp.tanner.ns, 0)
    direct = _cabi.Plan.compile(prob)
    _cabi.Lowered(prob).save(tmp_path / "d7.tqlw")
    loaded = _cabi.Plan.from_lowered(_cabi.Lowered.load(tmp_path / "d7.tqlw"))
    assert loaded.query(_cabi.Q_SWEEP) == direct.query(_cabi.Q_SWEEP) == 1
    words = tq.pack_bits(np.concatenate([sx, sz], axis=1))
    c0, l0 = direct.decode_map(words)
    c1, l1 = loaded.decode_map(words)
    assert np.array_equal(c0, c1) and np.array_equal(l0, l1)
    t5, em5 = _css_case(tq, tq.SurfaceCode(5, 5))
    _, _, f5, ch5, dims, _, _, _ = D._tnmmap_css_graph(tq.get_problem(t5, em5))
    prob5 = _cabi.Problem(f5, ch5, S.SUMPROD, dims[0], dims[1], dims[2], table_bits=-1)
    _cabi.Lowered(prob5).save(tmp_path / "m5.tqlw")
    a, b = _cabi.Plan.compile(prob5), _cabi.Plan.from_lowered(_cabi.Lowered.load(tmp_path / "m5.tqlw"))
    _, _, sx5, sz5 = _syndromes(t5, em5, 6, 300)
    w5 = tq.pack_bits(np.concatenate([sx5, sz5], axis=1))
    ma, mb = a.decode_marginal(w5), b.decode_marginal(w5)
    assert np.array_equal(ma[0], mb[0]) and np.array_equal(ma[1], mb[1])


def test_library_communicator_single_rank(tq):
    """tqec_comm_* with one rank: the NCCL all-reduce inside the fused Monte-Carlo pipeline leaves the counters as they
    are, and the stand-alone all-reduce returns its input (the 2 / 4 / 8-rank case runs under torchrun in bench.py)."""
    from tensorqec.jl_b200 import _cabi
    t, em = _css_case(tq, tq.SurfaceCode(5, 5))
    mc = tq.MonteCarlo(t, tq.TNMAP(), em)
    comm = _cabi.Comm(1, 0, _cabi.Comm.unique_id(), 0)
    a, _ = mc.run(5000, seed=3)
    b, _ = mc.run(5000, seed=3, comm=comm)
    assert np.array_equal(a, b) and b[3] == 5000
    assert np.array_equal(comm.allreduce_counts([1, 2, 3, 4]), [1, 2, 3, 4])
    comm.close()


def test_dem_d5_circuit_level_against_golden(tq, golden_dir):
    """BASELINE configs[3] at full size: d = 5 x 5 rounds rotated-surface memory circuit with circuit-level noise
    (1605 mechanisms, 120 detectors, 29-bit frontier) -> detector_error_model -> compile(TNMMAP) -> global-memory executor.
    16 detector patterns against marginals the CPU oracle computed along the REVERSED absorption order
    (tests/golden/make_dem_d5_golden.py; ~1 minute per shot on 8 cores), rtol 1e-10."""
    import json
    from tensorqec.jl_b200 import _cabi
    g = json.load(open(golden_dir / "dem_d5_r5_golden.json"))
    txt = tq.surface_memory_circuit(5, 5, "Z", 1e-3, 1e-3, 1e-3, 1e-3)
    dem = tq.detector_error_model(tq.parse_stim_string(txt))
    assert len(dem.error_rates) == g["n_mechanisms"] and dem.n_detectors == g["n_detectors"]
    ct = tq.compile(tq.TNMMAP(), dem)
    assert ct.plan.query(_cabi.Q_WIDE) == 1 and ct.plan.lowered["w_cap"] <= 29
    B = len(g["detectors"])
    syn = np.zeros((B, dem.n_detectors), dtype=np.uint8)
    for b, dets in enumerate(g["detectors"]):
        syn[b, dets] = 1
    ref = np.array([[float.fromhex(x) for x in row] for row in g["marginal"]])
    res = tq.decode(ct, tq.SimpleSyndrome(syn))
    got = res.marginal.reshape(B, -1, order="F")
    assert np.allclose(got, ref, rtol=MAR_RTOL, atol=0)
    assert np.array_equal(res.sector, ref.argmax(axis=1))
    assert (res.sector == np.array(g["true_observable"])).mean() >= 0.8
    assert tq.SimpleSyndrome(syn) == tq.syndrome_extraction(res.error_pattern, ct.tanner) and np.all(res.success_tag)


def test_dynamic_rescaling_survives_fp64_underflow(tq):
    """TNMMAP, d = 9, p = 1e-20 per Pauli, syndromes of weight ~40 and the all-ones syndrome (at least 20 errors): their
    probabilities lie far below the FP64 range (1e-400 and less), so the statically scaled plan returns zeros.  With
    TNMMAP(dynamic_rescale=True) the global-memory executor carries an int32 exponent per shot: mantissas and exponents
    match the recurrence oracle run with per-step renormalisation (rtol 1e-10), and the decoded sector is the argmax."""
    d = 9
    t = tq.CSSTannerGraph(tq.SurfaceCode(d, d))
    em = tq.iid_error(1e-20, t)
    rng = np.random.default_rng(12)
    syn = rng.integers(0, 2, size=(6, 80), dtype=np.uint8)
    syn[0] = 1
    syn[1] = 0
    s = tq.CSSSyndrome(syn[:, :40], syn[:, 40:])
    static = tq.compile(tq.TNMMAP(table_bits=0), t, em)
    r0 = tq.decode(static, s)
    assert (r0.marginal.reshape(6, -1)[0] == 0).all() and r0.marginal.reshape(6, -1)[1].max() > 0.99   # underflow vs the trivial syndrome
    ct = tq.compile(tq.TNMMAP(table_bits=0, dynamic_rescale=True), t, em)
    from tensorqec.jl_b200 import _cabi
    assert ct.plan.query(_cabi.Q_WIDE) == 1
    mant, lg, arg = ct.plan.decode_marginal_log2(tq.pack_bits(syn))
    sch = ct.schedule
    ref_m, ref_e = frontier.run(sch.factors, sch.checks, sch.order, 1, syn, sch.n_vars, rescale=True)
    assert (mant.max(axis=1) > 0).all() and lg[0] < -1200 and lg[1] > -60
    with np.errstate(invalid="ignore", divide="ignore"):
        ratio = (mant / ref_m) * np.exp2((lg.astype(np.int64) - ref_e)[:, None].astype(np.float64))
    assert np.allclose(ratio[ref_m > 0], 1.0, rtol=MAR_RTOL, atol=0)
    assert np.array_equal(arg, ref_m.argmax(axis=1))
    # through decode(): sectors are right although the probabilities themselves flush to zero
    r1 = tq.decode(ct, s)
    assert np.array_equal(r1.sector, arg) and tq.syndrome_extraction(r1.error_pattern, t) == s


def test_dynamic_rescaling_on_butterfly_passes(tq):
    """The same on the butterfly executor (k_wide_bf): phenomenological d = 3 x 3 rounds DEM with every mechanism at
    p = 1e-60; detector patterns that need seven mechanisms have probabilities near 1e-420.  The lowering ends
    a pass before its ratios r = p / (1 - p) can take a shot below the FP64 range, the kernel rescales between passes."""
    import os
    from tensorqec.jl_b200 import _cabi
    from tensorqec.jl_b200.dem import DetectorErrorModel
    dem0 = tq.parse_dem_file(os.path.join(os.path.dirname(__file__), "golden", "surface_d3_r3_phenom.dem"))
    dem = DetectorErrorModel([1e-60] * len(dem0.error_rates), dem0.flipped_detectors, dem0.detector_list, dem0.logical_list)
    ct = tq.compile(tq.TNMMAP(table_bits=0, dynamic_rescale=True), dem)
    assert ct.plan.query(_cabi.Q_WIDE) == 1
    nd = dem.n_detectors
    rng = np.random.default_rng(21)
    ep = (rng.random((6, len(dem.error_rates))) < 0.2).astype(np.uint8)
    syn = (ep @ ct.tanner.H.T.astype(np.int64) % 2).astype(np.uint8)
    syn[1] = 0
    assert syn.shape[1] == nd
    mant, lg, arg = ct.plan.decode_marginal_log2(tq.pack_bits(syn))
    sch = ct.schedule
    ref_m, ref_e = frontier.run(sch.factors, sch.checks, sch.order, 1, syn, sch.n_vars, rescale=True)
    assert (mant.max(axis=1) > 0).all() and lg.min() < -1100 and lg[1] > -60
    with np.errstate(invalid="ignore", divide="ignore"):
        ratio = (mant / ref_m) * np.exp2((lg.astype(np.int64) - ref_e)[:, None].astype(np.float64))
    assert np.allclose(ratio[ref_m > 0], 1.0, rtol=MAR_RTOL, atol=0)
    assert np.array_equal(arg, ref_m.argmax(axis=1))


def test_table_decoder_batched_lookup(tq):
    """TableDecoder (truthtable.jl) on the GPU: d = 5 surface code, all errors up to weight 2 tabulated; every sampled
    syndrome found in the table decodes to the tabulated pattern (which reproduces the syndrome), the others report
    success_tag False; d = 9 exercises two-word keys."""
    for d, w in ((5, 2), (9, 1)):
        t, em = _css_case(tq, tq.SurfaceCode(d, d), p=0.01)
        ct = tq.compile(tq.TableDecoder(w), t, em)
        ex, ez, sx, sz = _syndromes(t, em, 3, 20000)
        res = tq.decode(ct, tq.CSSSyndrome(sx, sz))
        ok = np.asarray(res.success_tag)
        assert 0.2 < ok.mean() <= 1.0
        rsx, rsz = gf2.css_syndrome(res.error_pattern.xerror, res.error_pattern.zerror, t.stgx.H, t.stgz.H)
        assert np.array_equal(rsx[ok], sx[ok]) and np.array_equal(rsz[ok], sz[ok])
        assert not res.error_pattern.xerror[~ok].any() and not res.error_pattern.zerror[~ok].any()
        # against a dictionary look-up on the host
        keys = {tuple(k): v for k, v in zip(ct.table.keys.tolist(), tq.unpack_bits(ct.table.values, 2 * d * d))}
        words = tq.pack_bits(np.concatenate([sx, sz], axis=1))
        for b in range(0, 20000, 97):
            v = keys.get(tuple(words[b].tolist()))
            assert (v is not None) == bool(ok[b])
            if v is not None:
                assert np.array_equal(v[:d * d], res.error_pattern.xerror[b]) and np.array_equal(v[d * d:], res.error_pattern.zerror[b])
        # weight <= w errors are always corrected exactly up to a stabilizer: the decoded pattern has the same syndrome and
        # (for w <= (d-1)/2) the same logical class
        lx, lz = tq.logical_operator(t)
        light = (ex | ez).sum(axis=1) <= w
        assert ok[light].all()
        fl = gf2.check_logical_error_css(ex[light], ez[light], res.error_pattern.xerror[light], res.error_pattern.zerror[light], lx, lz)
        assert not fl.any()


@pytest.mark.parametrize("code", ["steane", "surface3"])
def test_encoder_network_inference(tq, code):
    """`syndrome_inference` on the encoder (Clifford) network -- labels of dimension 4 as bit pairs, gate
    tensors as parity relations + sign tables -- through the sum-product executor (Steane: 14-bit frontier; 9-qubit
    surface code: 18 bits on the global-memory executor), batched over syndromes, against brute-force enumeration of the
    reference's network (all 4^n Pauli strings; rtol 1e-10).  Then the reference's own test flow in the Pauli frame
    (test/decoding/inferenceswithencoder.jl:80-130): a single-qubit error -> measured syndrome -> inference ->
    correction clears the syndrome and leaves no logical operator."""
    from oracle import encoder_bruteforce as bf
    from tensorqec.jl_b200 import encoder as E
    t = tq.CSSTannerGraph(tq.SteaneCode() if code == "steane" else tq.SurfaceCode(3, 3))
    n = t.stgx.nq
    qc, data, b = E.encode_stabilizers(t)
    measured = sorted(b.ordering[: b.matrix.shape[0]])
    p = [[0.85, 0.05, 0.05, 0.05]] * n
    ci = E.CompiledInference(qc, n, p, measured)
    rng = np.random.default_rng(8)
    syn = rng.integers(0, 2, size=(12, len(measured)), dtype=np.uint8)
    syn[0] = 0
    mar = ci.marginals(syn)
    for bidx in (0, 1, 5, 11):
        ref = bf.marginals(qc, n, p, {q: int(syn[bidx, i]) for i, q in enumerate(measured)})
        for k in range(n):
            assert np.allclose(mar[k][bidx], ref[k], rtol=MAR_RTOL, atol=1e-14), (k, bidx)
    # the reference's flow: X (then Z, Y) on one qubit, syndromes from commutation with the code's generators
    lx, lz = tq.logical_operator(t)
    for q, pauli in ((3, 1), (n - 1, 3), (1, 2)):
        ex = np.zeros(n, dtype=np.uint8)
        ez = np.zeros(n, dtype=np.uint8)
        ex[q], ez[q] = pauli in (1, 2), pauli in (2, 3)
        sx, sz = gf2.css_syndrome(ex[None], ez[None], t.stgx.H, t.stgz.H)
        outcome = 1 - 2 * np.concatenate([sx[0], sz[0]]).astype(int)            # +1 / -1 per generator (X-type first)
        corr = E.inference(outcome, b, qc, p)
        cx = np.array([int(a in (1, 2)) for a in corr], dtype=np.uint8)
        cz = np.array([int(a in (2, 3)) for a in corr], dtype=np.uint8)
        rsx, rsz = gf2.css_syndrome((ex ^ cx)[None], (ez ^ cz)[None], t.stgx.H, t.stgz.H)
        assert not rsx.any() and not rsz.any()
        assert not gf2.check_logical_error_css(ex[None], ez[None], cx[None], cz[None], lx, lz).any()


def test_bp_osd_decoder(tq):
    """BPDecoder (bposd.jl) batched on the GPU against the numpy restatement of the reference's loops: the reference's own
    example (test/decoding/bposd.jl:5-14), then the Z-check graph of the d = 5 surface code at p = 5 %: every returned
    pattern reproduces its syndrome (OSD guarantees it), the BP-converged flags and the patterns agree with the oracle
    (the tanh / atanh of the two implementations may differ in the last bit, hence a 99 % bar instead of equality)."""
    from oracle import bposd as obp
    tanner = tq.SimpleTannerGraph(7, [[0, 1, 2, 3], [1, 2, 4, 6], [2, 3, 4, 5]])
    ct = tq.compile(tq.BPDecoder(), tanner)
    e0 = np.array([1, 0, 0, 0, 0, 0, 0], dtype=np.uint8)
    res = tq.decode(ct, tq.syndrome_extraction(e0, tanner))
    assert res.success_tag and np.array_equal(res.error_pattern, e0)
    t = tq.CSSTannerGraph(tq.SurfaceCode(5, 5)).stgz
    em = tq.iid_error(0.05, 25)
    ct = tq.compile(tq.BPDecoder(), t, em)
    ep = tq.random_error_pattern(em, seed=4, shots=3000)
    syn = tq.syndrome_extraction(ep, t)
    res = tq.decode(ct, syn)
    assert np.all(res.success_tag)
    assert tq.syndrome_extraction(res.error_pattern, t) == syn
    nobp = tq.decode(tq.compile(tq.BPDecoder(100, False), t, em), syn)
    conv = np.asarray(nobp.success_tag)
    assert 0.5 < conv.mean() < 1.0 and not nobp.error_pattern[~conv].any()
    H = t.H.astype(np.uint8)
    agree = same_flag = 0
    for b in range(300):
        ok, e, by_bp = obp.decode(H, t.s2q, t.q2s, em.p, syn.s[b])
        agree += np.array_equal(e, res.error_pattern[b])
        same_flag += bool(by_bp) == bool(conv[b])
    assert agree >= 297 and same_flag >= 297


def test_multi_round_qec_accepts_other_decoders(tq):
    """threshold.jl:1-19 takes any decoder: TNMAP runs fused, TNMMAP / TableDecoder run the four stages as separate batched
    calls on the same Philox shots, so the TNMAP and TNMMAP rates are close and the shots are identical."""
    t = tq.CSSTannerGraph(tq.SurfaceCode(3, 3))
    em = tq.iid_error(0.03, t)
    a = tq.multi_round_qec(t, tq.TNMAP(), em, rounds=20000, seed=5)
    b = tq.multi_round_qec(t, tq.TNMMAP(), em, rounds=20000, seed=5)
    c = tq.multi_round_qec(t, tq.TableDecoder(2), em, rounds=20000, seed=5)
    assert 0 < a[2] < 0.1 and abs(a[2] - b[2]) < 0.01 and c[2] >= a[2] - 0.01
    r = tq.multi_round_qec(t, tq.TNMAP(), em, rounds=20000, seed=5, reference_prior=True)
    assert abs(r[2] - a[2]) < 0.02                               # decoding with the default 5 % prior instead of 3 %


def test_property_full_size_d9(tq):
    """BASELINE config 3 shape (d=9, p=0.05) at a size the oracle cannot follow shot by shot: size-independent
    properties -- every correction reproduces its syndrome, decoding is idempotent on its own output's syndrome,
    and the fused pipeline's counters equal the counters recomputed from the unfused stages."""
    d, B = 9, 200_000
    t, em = _css_case(tq, tq.SurfaceCode(d, d))
    mc = tq.MonteCarlo(t, tq.TNMAP(), em)
    counts, _ = mc.run(B, seed=9)
    ep = tq.random_error_pattern(em, seed=9, shots=B)
    syn = tq.syndrome_extraction(ep, t)
    res = tq.decode(mc.compiled, syn)
    assert tq.syndrome_extraction(res.error_pattern, t) == syn
    lx, lz = tq.logical_operator(t)
    fl = tq.check_logical_error(ep, res.error_pattern, lx, lz)
    assert counts[2] == int(fl.sum()) and counts[3] == B
    res2 = tq.decode(mc.compiled, tq.syndrome_extraction(res.error_pattern, t))
    assert np.array_equal(res2.error_pattern.xerror, res.error_pattern.xerror)
    assert np.array_equal(res2.logp, res.logp)
    # the MAP weight is at least the weight of the true error
    T = np.log(np.array([[0.85, 0.05], [0.05, 0.05]]))
    assert (res.logp >= T[ep.xerror, ep.zerror].sum(axis=1) - 1e-9).all()


def test_error_behaviour(tq):
    t = tq.CSSTannerGraph(tq.SurfaceCode(3, 3))
    ct = tq.compile(tq.TNMAP(), t)
    with pytest.raises(ValueError):
        tq.decode(ct, tq.CSSSyndrome(np.zeros(3, dtype=np.uint8), np.zeros(4, dtype=np.uint8)))
    with pytest.raises(TypeError):
        tq.decode(ct, tq.SimpleSyndrome(np.zeros(8, dtype=np.uint8)))
    with pytest.raises(tq.TqecError):
        tq.compile(tq.TNMAP(device=99), t)
    empty = tq.decode(ct, tq.CSSSyndrome(np.zeros((0, 4), dtype=np.uint8), np.zeros((0, 4), dtype=np.uint8)))
    assert empty.error_pattern.xerror.shape == (0, 9)


def test_cabi_error_codes_and_queries(tq):
    """The raw C ABI: error codes + messages on misuse, plan queries, device-pointer entry points."""
    import ctypes as C
    import torch
    from tensorqec.jl_b200 import _cabi
    lib = _cabi.lib()
    t = tq.CSSTannerGraph(tq.SurfaceCode(3, 3))
    ct = tq.compile(tq.TNMAP(), t)
    plan = ct.cd.plan
    # wrong semiring for the entry point
    out = np.zeros(4)
    rc = lib.tqec_decode_marginal(plan.h, out.ctypes.data_as(C.c_void_p), 1, out.ctypes.data_as(C.c_void_p), None)
    assert rc == -1 and b"sum-product" in lib.tqec_last_error()
    # NULL buffers
    assert lib.tqec_decode_map(plan.h, None, 5, None, None) == -1
    assert lib.tqec_decode_map(plan.h, None, 0, None, None) == 0          # empty batch is fine
    # malformed schedule: widths that do not chain
    sch = ct.cd.schedule
    bad = sch.hdr.copy()
    bad[1, 1] += 1
    hdr = np.ascontiguousarray(bad, dtype=np.int32)
    ints = np.ascontiguousarray(sch.ints, dtype=np.int32)
    tabs = np.ascontiguousarray(sch.tables, dtype=np.float64)
    obs = np.zeros(1, dtype=np.int32)
    d = _cabi.PlanDesc(0, sch.n_vars, sch.n_checks, 0, len(sch.steps), sch.w_max, hdr.ctypes.data_as(C.POINTER(C.c_int32)),
                       ints.ctypes.data_as(C.POINTER(C.c_int32)), ints.size, tabs.ctypes.data_as(C.POINTER(C.c_double)), tabs.size,
                       obs.ctypes.data_as(C.POINTER(C.c_int32)), 0)
    h = C.c_void_p()
    assert lib.tqec_plan_create(C.byref(d), C.byref(h)) == -1 and not h.value
    assert b"w_in" in lib.tqec_last_error()
    # queries
    g = plan.geometry()
    assert g["team_threads"] == 32 and g["sm_count"] >= 100 and g["candidates_per_shot"] == 104
    n0 = plan.query(_cabi.Q_LAUNCHES)
    # device-pointer entry point on torch memory, asynchronous on the current stream
    syn = ((np.arange(256)[:, None] >> np.arange(8)) & 1).astype(np.uint8)
    words = tq.pack_bits(syn)
    d_syn = torch.from_numpy(words.view(np.int64)).cuda()
    d_cor = torch.zeros((256, plan.ncw), dtype=torch.int64, device="cuda")
    d_lp = torch.zeros(256, dtype=torch.float64, device="cuda")
    plan.decode_map_dev(d_syn.data_ptr(), 256, d_cor.data_ptr(), d_lp.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    corr, lp = plan.decode_map(words)
    assert np.array_equal(d_cor.cpu().numpy().view(np.uint64), corr) and np.array_equal(d_lp.cpu().numpy(), lp)
    assert plan.query(_cabi.Q_LAUNCHES) == n0 + 2
    # GF(2) shape errors surface as Python exceptions before the ABI is reached
    with pytest.raises(ValueError):
        tq.syndrome_extraction(np.zeros(5, dtype=np.uint8), t.stgz)
