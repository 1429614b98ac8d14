"""CPU tests of the global-memory lowering (tensorqec.jl_b200/wide.py): the flat pass / step tables, executed by the
table-level emulator exactly as `k_wide_pass` reads them, against the recurrence oracle written over named axes."""
import numpy as np
import pytest

import tensorqec.jl_b200 as tq
from oracle import frontier, wide_emulator
from tensorqec.jl_b200 import decoding, schedule as S, wide as W


def _dem_graph(dem):
    tanner = tq.dem2tanner(dem)
    ne, nd = tanner.nq, tanner.ns
    l2q = [[e for e in range(ne) if l in dem.flipped_detectors[e]] for l in dem.logical_list]
    factors = [S.Factor((e,), np.array([1.0 - p, p])) for e, p in enumerate(dem.error_rates)]
    checks = [S.Check(tuple(c), "syn", d) for d, c in enumerate(tanner.s2q)]
    checks += [S.Check(tuple(c), "obs", l) for l, c in enumerate(l2q)]
    return factors, checks, ne, nd, len(l2q), tanner


@pytest.mark.parametrize("t_max,low", [(8, 4), (8, 2), (12, 4)])
def test_wide_tables_match_recurrence_circuit_level_d3(t_max, low):
    txt = tq.surface_memory_circuit(3, 3, "Z", 0.01, 0.01, 0.01, 0.01)
    dem = tq.detector_error_model(tq.parse_stim_string(txt))
    factors, checks, ne, nd, no, tanner = _dem_graph(dem)
    wp = W.lower_wide(factors, checks, S.SUMPROD, ne, nd, no, t_max=t_max, low_bits=low)
    assert all(p.t_peak <= t_max for p in wp.passes)
    assert sum(len(p.steps) for p in wp.passes) == ne
    if t_max == 8:
        assert any(p.w_in > p.t_in for p in wp.passes), "no pass with spectator bits: the tile path is not exercised"
    rng = np.random.RandomState(3)
    e = (rng.rand(6, ne) < np.array(dem.error_rates) * 8).astype(np.uint8)
    syn = (e @ tanner.H.T.astype(np.int64)) % 2
    got = wide_emulator.run(wp, syn)
    ref = frontier.run(wp.factors, wp.checks, wp.order, 1, syn, ne)
    assert np.allclose(got, ref, rtol=1e-13, atol=0)


def test_wide_tables_css_tnmmap_rank2_factors():
    """CSS marginal network (rank-2 priors, 2 open observables, up to 4 candidates per output) through the wide lowering."""
    t = tq.CSSTannerGraph(tq.SurfaceCode(5, 5))
    em = tq.iid_error(0.03, t)
    prob = decoding.IndependentDepolarizingDecodingProblem(t, em)
    n = 25
    lx, lz = tq.logical_operator(t)
    factors = [S.Factor((i, i + n), S.flat_table(decoding.single_qubit_tensor(em.px[i], em.py[i], em.pz[i]))) for i in range(n)]
    checks = [S.Check(tuple(q + n for q in c), "syn", i) for i, c in enumerate(t.stgx.s2q)]
    checks += [S.Check(tuple(c), "syn", 12 + i) for i, c in enumerate(t.stgz.s2q)]
    checks += [S.Check(tuple(int(q) + n for q in np.flatnonzero(lx[0])), "obs", 0)]
    checks += [S.Check(tuple(int(q) for q in np.flatnonzero(lz[0])), "obs", 1)]
    wp = W.lower_wide(factors, checks, S.SUMPROD, 2 * n, 24, 2, t_max=5, low_bits=1)
    assert len(wp.passes) > 3
    rng = np.random.RandomState(5)
    syn = rng.randint(0, 2, size=(5, 24)).astype(np.uint8)
    got = wide_emulator.run(wp, syn)
    ref = frontier.run(wp.factors, wp.checks, wp.order, 1, syn, 2 * n)
    assert np.allclose(got, ref, rtol=1e-13, atol=0)


def test_spectral_order_narrows_the_3d_detector_graph():
    """d = 5 x 5 rounds circuit-level memory (BASELINE configs[3]): the spectral sweep finds a 29-bit front."""
    txt = tq.surface_memory_circuit(5, 5, "Z", 0.001, 0.001, 0.001, 0.001)
    dem = tq.detector_error_model(tq.parse_stim_string(txt))
    factors, checks, ne, nd, no, _ = _dem_graph(dem)
    sim = S._Sim(factors, checks)
    best = min(S._evaluate(o, sim)[0] for o in S.spectral_orders(sim))
    assert best <= 29
    wp = W.lower_wide(factors, checks, S.SUMPROD, ne, nd, no, order=min(S.spectral_orders(sim), key=lambda o: S._evaluate(o, sim)))
    assert wp.w_cap <= 29 and wp.bytes_per_shot < 1.2e11


@pytest.mark.parametrize("t_max", [12, 9, 7])
def test_butterfly_encoding_of_the_library_lowering_matches_recurrence(t_max):
    """The library's own lowering (C++) adds a butterfly encoding to every pass made of rank-1 factors (k_wide_bf): fixed
    positions, closed checks folded into the addresses, cosets of 2^5 entries, ratios r = t1 / t0 with the product of the
    t0 in the output scale.  The emulator executes those tables the way the kernel does; against the recurrence oracle
    and against the same plan's generic tables (k_wide_pass)."""
    from tensorqec.jl_b200 import _cabi
    txt = tq.surface_memory_circuit(3, 3, "Z", 0.01, 0.01, 0.01, 0.01)
    dem = tq.detector_error_model(tq.parse_stim_string(txt))
    factors, checks, ne, nd, no, tanner = _dem_graph(dem)
    lw = _cabi.Lowered(_cabi.Problem(factors, checks, S.SUMPROD, ne, nd, no, flags=_cabi.COMPILE_FORCE_WIDE, wide_t_max=t_max))
    off = lw.get(_cabi.LW_WD_BF_OFF)
    assert len(off) == lw.meta["n_pass"] and (off >= 0).all(), "every pass of a detector error model is a butterfly pass"
    bi = lw.get(_cabi.LW_WD_BF_INTS)
    if t_max == 7:
        zs = [int(bi[o + 20 + 16 * g + 10]) for o in off for g in range(int(bi[o]))]
        assert any(zs), "no position is reused: the zeroing of reopened positions is not exercised"
    rng = np.random.RandomState(7)
    e = (rng.rand(6, ne) < np.array(dem.error_rates) * 8).astype(np.uint8)
    syn = (e @ tanner.H.T.astype(np.int64)) % 2
    got = wide_emulator.run_lowered(lw, no, syn, butterfly=True)
    generic = wide_emulator.run_lowered(lw, no, syn, butterfly=False)
    order = [int(i) for i in lw.get(_cabi.LW_ORDER)]
    merged = S.merge_overlapping(list(factors), ne, {v for c in checks for v in c.vars}, allow_negative=True)
    ref = frontier.run(merged, checks, order, 1, syn, ne)
    assert np.allclose(generic, ref, rtol=1e-13, atol=0)
    assert np.allclose(got, ref, rtol=1e-12, atol=0)


@pytest.mark.parametrize("seed", range(6))
def test_butterfly_encoding_on_random_hypergraphs(seed):
    """Random detector error models (mechanisms flipping 1-4 of 14-18 detectors and sometimes an observable; repeated
    masks, mechanisms that open several detectors at once, detectors touched once) lowered with small tiles, so that
    positions are reused, several checks open in one step and passes have many groups.  The butterfly tables executed by
    the emulator against the recurrence oracle."""
    from tensorqec.jl_b200 import _cabi
    rng = np.random.RandomState(100 + seed)
    nd, nm, no = 14 + seed % 5, 40 + 6 * seed, 1 + seed % 2
    masks = []
    for _ in range(nm):
        k = rng.randint(1, 5)
        m = sorted(rng.choice(nd, size=k, replace=False).tolist())
        masks.append(m)
    masks += masks[:5]                                           # repeated masks: dependent steps with one coordinate
    p = rng.uniform(0.001, 0.3, size=len(masks))
    obs = [sorted(set(rng.choice(len(masks), size=6, replace=False).tolist())) for _ in range(no)]
    factors = [S.Factor((e,), np.array([1.0 - p[e], p[e]])) for e in range(len(masks))]
    checks = [S.Check(tuple(e for e, m in enumerate(masks) if d in m), "syn", d) for d in range(nd)]
    used = [d for d in range(nd) if checks[d].vars]
    checks = [S.Check(checks[d].vars, "syn", i) for i, d in enumerate(used)]
    nd = len(used)
    checks += [S.Check(tuple(o), "obs", l) for l, o in enumerate(obs)]
    ne = len(masks)
    t_max = 6 + seed % 3
    lw = _cabi.Lowered(_cabi.Problem(factors, checks, S.SUMPROD, ne, nd, no, flags=_cabi.COMPILE_FORCE_WIDE, wide_t_max=t_max))
    off = lw.get(_cabi.LW_WD_BF_OFF)
    assert (off >= 0).sum() >= len(off) - 1, "rank-1 passes should get a butterfly block"
    H = np.zeros((nd, ne), dtype=np.int64)
    for c in checks[:nd]:
        H[c.index, list(c.vars)] = 1
    e = (rng.rand(5, ne) < 0.15).astype(np.int64)
    syn = ((e @ H.T) % 2).astype(np.uint8)
    got = wide_emulator.run_lowered(lw, no, syn, butterfly=True)
    order = [int(i) for i in lw.get(_cabi.LW_ORDER)]
    ref = frontier.run(factors, checks, order, 1, syn, ne)
    assert np.allclose(got, ref, rtol=1e-11, atol=0)
