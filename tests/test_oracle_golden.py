"""Pins the oracle (and the host data model it is built on) against every known-answer value the reference's own
tests hold for this path.  CPU only."""
import os

import numpy as np
import pytest

from oracle import bruteforce, cref, dense, emulator, frontier, gf2, networks, philox

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_parity_check_matrix_entries():
    # test/decoding/tndecoder.jl:8-14 (1-based Julia indices -> 0-based)
    m = networks.parity_check_matrix(4)
    assert m.shape == (2,) * 5
    assert m[0, 0, 0, 0, 0] == 1
    assert m[0, 0, 0, 0, 1] == 0
    assert m[0, 1, 0, 1, 0] == 1
    assert m[1, 0, 1, 1, 0] == 0


def test_tnmmap_marginal_golden_dense_and_frontier(tq):
    # test/decoding/tndecoder.jl:101-110, atol 1e-10
    gold = np.array([[0.3972875040000002, 0.004284496000000001], [0.0, 0.0]])
    t = tq.CSSTannerGraph(tq.SurfaceCode(3, 3))
    lx, lz = tq.logical_operator(t)
    p, z = np.full(9, 0.1), np.zeros(9)
    net = networks.tnmmap_css_network(t, lx, lz, p, z, z)
    assert np.allclose(dense.contract_sumproduct(net), gold, atol=1e-10)
    assert np.allclose(cref.DensePlan(net, 8, 18, False).run(np.zeros((1, 8), dtype=np.uint8)).reshape(2, 2, order="F"), gold, atol=1e-10)
    _, _, sch, *_ = tq.tnmmap_css_schedule(tq.TNMMAP(), tq.IndependentDepolarizingDecodingProblem(t, tq.IndependentDepolarizingError(p, z, z)))
    zero = np.zeros((1, 8), dtype=np.uint8)
    for got in (frontier.run(sch.factors, sch.checks, sch.order, 1, zero, 18), emulator.run(sch, zero), cref.FrontierPlan(sch).run(zero)):
        assert np.allclose(got.reshape(2, 2, order="F"), gold, atol=1e-10)
    # brute force over the 2^9 X patterns
    nq, s2q, pix, pri = networks.general_problem_css(t, p, z, z)
    en = bruteforce.Enumeration(nq, s2q, pix, pri)
    Lg = np.zeros((2, 18), dtype=np.uint8)
    Lg[0, 9:], Lg[1, :9] = lx[0], lz[0]
    assert np.allclose(en.marginal(np.zeros(8, dtype=np.uint8), Lg).reshape(2, 2, order="F"), gold, atol=1e-10)


def test_logical_operator_golden(tq):
    # test/codes/code_distance.jl:54-59
    lx, lz = tq.logical_operator(tq.CSSTannerGraph(tq.SurfaceCode(3, 3)))
    assert lx.tolist() == [[0, 0, 0, 0, 0, 0, 1, 1, 1]]
    assert lz.tolist() == [[1, 0, 0, 0, 1, 0, 0, 0, 1]]
    # lx = last row of the lattice, lz = main diagonal for d = 5, 7, 9
    for d in (5, 7, 9):
        t = tq.CSSTannerGraph(tq.SurfaceCode(d, d))
        lx, lz = tq.logical_operator(t)
        assert np.flatnonzero(lx[0]).tolist() == list(range(d * (d - 1), d * d))
        assert np.flatnonzero(lz[0]).tolist() == [i * (d + 1) for i in range(d)]
        assert not ((t.stgz.H @ lx.T) & 1).any() and not ((t.stgx.H @ lz.T) & 1).any()
        assert int(lx[0] @ lz[0]) & 1 == 1


def test_null_space_golden(tq):
    # test/codes/code_distance.jl:40-52
    H = np.array([[0, 0, 0, 1, 1, 1, 1], [0, 1, 1, 0, 0, 1, 1], [1, 0, 1, 0, 1, 0, 1]], dtype=np.uint8)
    ker = tq.null_space(H)
    assert ker.shape == (4, 7)
    assert not ((H @ ker.T) & 1).any()
    assert all(r.any() for r in ker)                          # the reference counts one pivot per row: rank 4
    from tensorqec.jl_b200.tanner import gf2_right_inverse
    assert gf2_right_inverse(ker)[1] == 4


def test_surface_code_layout(tq):
    # src/codes/codes.jl:14-26 comment (1-based): X 36, 1245, 5689, 47 ; Z 12, 2356, 4578, 89 ; generation order per SURVEY A.1
    t = tq.CSSTannerGraph(tq.SurfaceCode(3, 3))
    assert t.stgx.s2q == [[0, 1, 3, 4], [4, 5, 7, 8], [2, 5], [3, 6]]
    assert t.stgz.s2q == [[1, 2, 4, 5], [3, 4, 6, 7], [0, 1], [7, 8]]
    for d, nchk, nb in ((5, 12, 4), (7, 24, 6), (9, 40, 8)):
        t = tq.CSSTannerGraph(tq.SurfaceCode(d, d))
        assert (t.stgx.ns, t.stgz.ns) == (nchk, nchk)
        assert sum(len(s) == 2 for s in t.stgx.s2q) == nb and sum(len(s) == 2 for s in t.stgz.s2q) == nb
        assert not ((t.stgx.H @ t.stgz.H.T) & 1).any()


def test_tanner_graph_golden(tq):
    # test/codes/ldpc.jl:7-23 (0-based)
    tg = tq.SimpleTannerGraph(5, [[0, 1, 2, 3], [1, 2, 3, 4]])
    assert tg.q2s == [[0], [0, 1], [0, 1], [0, 1], [1]]
    assert tg.s2q == [[0, 1, 2, 3], [1, 2, 3, 4]] and tg.ns == 2
    assert tg.H.tolist() == [[1, 1, 1, 1, 0], [0, 1, 1, 1, 1]]
    tg2 = tq.SimpleTannerGraph(H=np.array([[1, 1, 1, 1, 0], [0, 1, 1, 1, 1]]))
    assert tg2.q2s == tg.q2s and tg2.s2q == tg.s2q
    assert tg.H.sum() == sum(map(len, tg.q2s)) == sum(map(len, tg.s2q))


def test_gf2_oracle_known_answers(tq):
    # test/codes/ldpc.jl:34-39 ; test/decoding/error_model.jl:21-33
    H = np.array([[1, 1, 1, 1, 0], [0, 1, 1, 1, 1]])
    assert gf2.syndrome_extraction([1, 0, 1, 1, 0], H).tolist() == [1, 0]
    lx, lz = tq.logical_operator(tq.CSSTannerGraph(tq.SurfaceCode(3, 3)))
    z9 = np.zeros(9, dtype=np.uint8)
    assert bool(gf2.check_logical_error(z9, [1, 1, 1, 0, 0, 0, 0, 0, 0], lz))
    assert not bool(gf2.check_logical_error(z9, [1, 1, 0, 1, 1, 0, 0, 0, 0], lz))
    assert not bool(gf2.check_logical_error_css(z9, z9, z9, z9, lx, lz))


def test_mod2_and_bitmul(tq):
    # test/codes/mod2.jl:3-31
    a, b = tq.Mod2(False), tq.Mod2(True)
    assert repr(a) == "0₂" and repr(b) == "1₂"
    assert a + b == tq.Mod2(True) and a + a == tq.Mod2(False) and b + b == tq.Mod2(False)
    assert a * b == tq.Mod2(False) and b * b == tq.Mod2(True)
    assert -b == tq.Mod2(True) and a - b == tq.Mod2(True) and b - a == tq.Mod2(True)
    assert a.iszero() and not b.iszero()
    rng = np.random.default_rng(0)
    A = rng.integers(0, 2, size=(300, 1000)).astype(np.uint8)
    B = rng.integers(0, 2, size=(1000, 200)).astype(np.uint8)
    assert np.array_equal(tq.bitmul(A, B), gf2.bitmul(A, B))
    w = tq.pack_bits(A)
    assert w.shape == (300, 16) and np.array_equal(tq.unpack_bits(w, 1000), A)
    assert int(w[0, 0]) & 1 == A[0, 0] and (int(w[0, 1]) >> 3) & 1 == A[0, 67]


def test_integer_thresholds_decide_like_the_float_compare():
    """The samplers compare the raw 53-bit draw with ceil(p * 2^53) instead of converting it to a double and comparing with p
    (k_sample_depol / k_sample_errors, csrc/tqec_gf2.cu): the same decision for every draw, because u = bits * 2^-53 and
    p * 2^53 are both exact.  Checked on random probabilities, on draws right at the threshold and on the corner values."""
    import math
    rng = np.random.default_rng(53)
    ps = np.concatenate([rng.uniform(0, 1, 2000), rng.uniform(0, 1e-9, 200), [0.0, 1.0, 0.5, 2.0 ** -53, 1 - 2.0 ** -53, 0.05 / 3]])
    for p in ps:
        thr = min(math.ceil(min(p, 1.0) * 2.0 ** 53), 2 ** 53) if p > 0 else 0
        near = [b for b in (thr - 2, thr - 1, thr, thr + 1, 0, 2 ** 53 - 1) if 0 <= b < 2 ** 53]
        bits = np.array(near + list(rng.integers(0, 2 ** 53, 20)), dtype=np.uint64)
        u = bits.astype(np.float64) * (1.0 / 9007199254740992.0)
        assert np.array_equal(u < p, bits < np.uint64(thr)), p
    # the sums of the depolarizing rule are rounded once, as doubles, before they become thresholds
    px, py, pz = 0.1, 0.2, 0.3
    assert math.ceil((px + py) * 2.0 ** 53) == math.ceil(float(np.float64(px) + np.float64(py)) * 2.0 ** 53)


def test_syndrome_containers_validate_once(tq):
    """SimpleSyndrome / CSSSyndrome check their bits when they are built and hand `decode` marked arrays (no second pass over a
    batch); views keep the mark, anything computed from them is a plain array again and would be checked."""
    from tensorqec.jl_b200.mod2 import ValidatedBits, as_bits
    sx = np.array([[0, 1, 1], [1, 0, 0]], dtype=np.uint8)
    s = tq.CSSSyndrome(sx, [[1, 0], [0, 1]])
    assert isinstance(s.sx, ValidatedBits) and isinstance(s.sz, ValidatedBits) and np.shares_memory(s.sx, sx)
    assert as_bits(s.sx) is s.sx and isinstance(s.sx[1:], ValidatedBits)
    for derived in (s.sx + 1, s.sx ^ s.sx, np.concatenate([s.sx, s.sz], axis=1), np.asarray(s.sx)):
        assert not isinstance(derived, ValidatedBits)
    with pytest.raises(ValueError):
        as_bits(s.sx + 1)
    with pytest.raises(ValueError):
        tq.SimpleSyndrome(np.array([0, 1, 2], dtype=np.uint8))
    with pytest.raises(ValueError):
        tq.CSSSyndrome(sx, [[0, 3]])
    assert tq.SimpleSyndrome([0, 1, 1]) == tq.SimpleSyndrome(np.array([0, 1, 1])) and s == tq.CSSSyndrome(sx.copy(), s.sz)
    assert int(s.sx.sum()) == 3 and bool((s.sx == sx).all())


def test_color488_and_steane(tq):
    # test/codes/codes.jl:176-181 (16 stabilizers) + SURVEY C.4 rows (1-based there)
    st = tq.stabilizers(tq.Color488(5))
    assert len(st) == 16
    H = tq.Color488(5).check_matrix()
    rows = [[int(c) + 1 for c in np.flatnonzero(r)] for r in H]
    assert rows == [[1, 2, 3, 4], [1, 3, 5, 6], [3, 4, 6, 7, 10, 11, 14, 15], [5, 6, 9, 10], [7, 8, 11, 12], [8, 12, 16, 17],
                    [9, 10, 13, 14], [11, 12, 15, 16]]
    assert not ((H @ H.T) & 1).any()
    t = tq.CSSTannerGraph(tq.Color488(5))
    lx, lz = tq.logical_operator(t)
    assert lx.shape == (1, 17) and lz.shape == (1, 17)
    # distance 5 (test/codes/codes.jl:180): the lightest non-trivial logical has weight 5 (exhaustive over 2^17)
    E = bruteforce.all_assignments(17)
    ok = ~((E @ H.T) & 1).any(axis=1) & (((E @ lz[0]) & 1) == 1)
    assert E[ok].sum(axis=1).min() == 5
    ts = tq.CSSTannerGraph(tq.SteaneCode())
    assert ts.stgx.s2q == [[0, 2, 4, 6], [1, 2, 5, 6], [3, 4, 5, 6]] == ts.stgz.s2q
    E = bruteforce.all_assignments(7)
    lxs, lzs = tq.logical_operator(ts)
    ok = ~((E @ ts.stgx.H.T) & 1).any(axis=1) & (((E @ lzs[0]) & 1) == 1)
    assert E[ok].sum(axis=1).min() == 3                     # test/codes/codes.jl:145-148


def test_dem_fixture(tq):
    # test/stim_parser/test_circuits/dem.dem: 21 mechanisms, 6 detectors, 1 observable
    dem = tq.parse_dem_file(os.path.join(GOLD, "dem.dem"))
    assert len(dem.error_rates) == 21 and dem.detector_list == list(range(6)) and dem.logical_list == [6]
    assert dem.error_rates[0] == 0.03946182850125499 and dem.flipped_detectors[0] == [0, 1, 2]
    assert dem.flipped_detectors[3] == [0, 1, 6]             # "D0 D1 L0"
    tg = tq.dem2tanner(dem)
    assert (tg.nq, tg.ns) == (21, 6)
    assert tg.s2q[0] == [0, 1, 2, 3, 4, 5, 6, 7]
    import pytest
    with pytest.raises(ValueError):
        tq.parse_dem_string("repeat 3 {\nerror(0.1) D0\n}")
    # union (not xor) of ^-separated components, non-error lines skipped
    d2 = tq.parse_dem_string("error(0.1) D0 D1 ^ D1 D2 L1\ndetector(1,1) D0\nlogical_observable L0\n")
    assert d2.flipped_detectors == [[0, 1, 2, 4]] and d2.logical_list == [3, 4]


def test_philox_known_answer():
    # Random123 known-answer vectors for Philox4x32-10 (kat_vectors): zero counter/key, and the pi-digits vector
    out = philox.philox4x32_10(0, 0, 0, 0, 0, 0)
    assert [int(x) for x in out] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    out = philox.philox4x32_10(0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF)
    assert [int(x) for x in out] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    out = philox.philox4x32_10(0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344, 0xA4093822, 0x299F31D0)
    assert [int(x) for x in out] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_sampling_rule_y_first():
    # error_model.jl:101-115: u < py -> Y, u < px+py -> X, u < px+py+pz -> Z
    u = philox.uniforms(5, 0, 4000, 3)
    ex, ez = philox.sample_depolarizing([0.05] * 3, [0.06] * 3, [0.1] * 3, 5, 0, 4000)
    assert np.array_equal(ex & ez, (u < 0.06).astype(np.uint8))
    assert np.array_equal(ex & (1 - ez), ((u >= 0.06) & (u < 0.05 + 0.06)).astype(np.uint8))
    assert np.array_equal(ez & (1 - ex), ((u >= 0.05 + 0.06) & (u < 0.05 + 0.06 + 0.1)).astype(np.uint8))
    assert abs(ex.mean() - 0.11) < 0.02 and abs(ez.mean() - 0.16) < 0.02   # test/decoding/error_model.jl:17-18
