"""The C++ lowering inside libtqec_cuda.so (csrc/tqec_lower*.cpp, entry points tqec_lower / tqec_plan_compile) against
the Python lowering it was ported from (schedule.py, sweep.py, wide.py), which stays in the repository as its oracle:
every emitted table must be IDENTICAL -- step headers, masks, factor tables, sweep records, traceback records, lane
tables, tabulated head states and configurations, wide pass / step headers.  Host code only: runs without a GPU."""
import numpy as np
import pytest

import tensorqec.jl_b200 as tq
from tensorqec.jl_b200 import _cabi, decoding as D, schedule as S, sweep as SW


def _same(a, b):
    a, b = np.asarray(a).reshape(-1), np.asarray(b).reshape(-1)
    if a.shape != b.shape:
        return False
    if a.dtype.kind == "f":
        return np.array_equal(a, b, equal_nan=True)
    return np.array_equal(a.astype(np.int64), b.astype(np.int64))


def _tnmap_graph(code, em=None):
    t = tq.CSSTannerGraph(code)
    em = em or tq.iid_error(0.05, t)
    gdp, _ = tq.reduce2general(t, em)
    return D._tnmap_graph(gdp) + (gdp.tanner.nq, gdp.tanner.ns)


def _dem_graph(dem):
    _, _, factors, checks, dims, _, _, _ = D._tnmmap_dem_graph(dem)
    return factors, checks, dims


def _check_schedule(lw, py):
    m = lw.meta
    sw = getattr(py, "sweep", None)
    assert m["kind"] == (1 if sw is not None else 0)
    assert _same(lw.get(_cabi.LW_ORDER), py.order)
    assert _same(lw.get(_cabi.LW_HDR), py.hdr) and _same(lw.get(_cabi.LW_INTS), py.ints)
    assert _same(lw.get(_cabi.LW_TABLES), py.tables)
    assert (m["n_steps"], m["w_max"], m["log2_scale"]) == (len(py.steps), py.w_max, py.log2_scale)
    if py.n_obs:
        assert _same(lw.get(_cabi.LW_OBS_SLOT), py.obs_slot)
    if sw is not None:
        for what, arr in [(_cabi.LW_SW_REC, sw.rec), (_cabi.LW_SW_TB, sw.tb), (_cabi.LW_SW_LANETAB, sw.lanetab),
                          (_cabi.LW_SW_TVALS, sw.tvals), (_cabi.LW_SW_HEAD_BITS, sw.head_bits),
                          (_cabi.LW_SW_HEAD_STATE, sw.head_state), (_cabi.LW_SW_OUT_INDEX, sw.out_index)]:
            assert _same(lw.get(what), arr), what
        if py.semiring == S.MAXPLUS:
            assert _same(lw.get(_cabi.LW_SW_HEAD_CFG), sw.head_cfg)
        assert (m["W"], m["sg"], m["n_ss"], m["bp_words"], m["head_steps"], m["conflicts"]) == \
            (sw.W, sw.sg, len(sw.ssteps), sw.bp_words, sw.head_steps, sw.conflicts)


@pytest.mark.parametrize("code", ["d3", "d5", "d7", "d9", "steane", "color488_5", "d4", "d6", "d8", "d5x7"])
def test_tnmap_tables_identical(code):
    c = {"steane": tq.SteaneCode(), "color488_5": tq.Color488(5), "d5x7": tq.SurfaceCode(5, 7)}.get(code) or tq.SurfaceCode(int(code[1:]), int(code[1:]))
    factors, checks, nq, ns = _tnmap_graph(c)
    py = D._tnmap_lower(tq.TNMAP(head_bits=12), factors, checks, nq, ns, None)
    lw = _cabi.Lowered(_cabi.Problem(factors, checks, S.MAXPLUS, nq, ns, 0, head_bits=12))
    _check_schedule(lw, py)
    if code == "d5":                                              # the shipped defaults agree too (head_bits = 0 -> the library's)
        _check_schedule(_cabi.Lowered(_cabi.Problem(factors, checks, S.MAXPLUS, nq, ns, 0)), D._tnmap_lower(tq.TNMAP(), factors, checks, nq, ns, None))
    if code in ("d5", "d7", "d9"):
        assert lw.meta["kind"] == 1, "odd-distance surface codes decode through the in-place patch sweep"


def test_tnmap_generic_noise_custom_order_and_head_bits():
    rng = np.random.default_rng(7)
    t = tq.CSSTannerGraph(tq.SurfaceCode(7, 7))
    em = tq.IndependentDepolarizingError(rng.uniform(0.01, 0.1, 49), rng.uniform(0.01, 0.1, 49), rng.uniform(0.01, 0.1, 49))
    factors, checks, nq, ns = _tnmap_graph(tq.SurfaceCode(7, 7), em)
    for head_bits in (6, 10):
        py = D._tnmap_lower(tq.TNMAP(head_bits=head_bits), factors, checks, nq, ns, None)
        _check_schedule(_cabi.Lowered(_cabi.Problem(factors, checks, S.MAXPLUS, nq, ns, 0, head_bits=head_bits)), py)
    order = list(range(48, -1, -1))                               # a caller-supplied order: general kernels, fused
    py = D._tnmap_lower(tq.TNMAP(head_bits=12), factors, checks, nq, ns, order)
    _check_schedule(_cabi.Lowered(_cabi.Problem(factors, checks, S.MAXPLUS, nq, ns, 0, order=order, head_bits=12)), py)


def test_overlapping_priors_are_merged_identically():
    """Correlated priors that share variables (general_decoding.jl accepts any SimpleTensorNetwork), a caller's order over
    ITS tensors (the order used to be misread as an order of the merged factors), a variable without prior."""
    rng = np.random.default_rng(3)
    ixs = [[0, 1], [1, 2], [3], [5, 4]]
    factors = [S.Factor(tuple(ix), rng.uniform(0.05, 1.0, 1 << len(ix))) for ix in ixs]
    checks = [S.Check((0, 2, 3), "syn", 0), S.Check((1, 3, 4), "syn", 1), S.Check((2, 5, 6), "syn", 2), S.Check((0, 6), "syn", 3)]
    for order in (None, [3, 0, 2, 1]):
        py = S.lower(factors, checks, S.MAXPLUS, 7, 4, 0, order=order)
        lw = _cabi.Lowered(_cabi.Problem(factors, checks, S.MAXPLUS, 7, 4, 0, order=order, flags=_cabi.COMPILE_NO_SWEEP))
        _check_schedule(lw, py)
        assert len(py.factors) == 4                               # {0,1,2} merged, {3}, {4,5}, unity factor for variable 6


@pytest.mark.parametrize("d", [3, 5, 7, 9, 6, (5, 7), (6, 8)])
def test_tnmmap_css_tables_identical(d):
    t = tq.CSSTannerGraph(tq.SurfaceCode(*d) if isinstance(d, tuple) else tq.SurfaceCode(d, d))
    _, _, factors, checks, dims, _, _, _ = D._tnmmap_css_graph(tq.get_problem(t, tq.iid_error(0.05, t)))
    py = D._sumprod_lower(tq.TNMMAP(), factors, checks, dims, None)
    _check_schedule(_cabi.Lowered(_cabi.Problem(factors, checks, S.SUMPROD, *dims)), py)


@pytest.mark.parametrize("name", ["dem.dem", "surface_d3_r3_phenom.dem", "surface_d5_r5_phenom.dem"])
def test_dem_tables_identical(name, golden_dir, monkeypatch):
    # the on-chip schedule (by default plans of rank-1 factors move to the global-memory executor from 10 bits on)
    monkeypatch.setenv("TQEC_SUMPROD_ONCHIP_WIDTH", "11")
    factors, checks, dims = _dem_graph(tq.parse_dem_file(str(golden_dir / name)))
    py = D._sumprod_lower(tq.TNMMAP(), factors, checks, dims, None)
    _check_schedule(_cabi.Lowered(_cabi.Problem(factors, checks, S.SUMPROD, *dims)), py)


def test_rank1_plans_of_ten_bits_and_more_go_to_the_butterfly_executor(golden_dir):
    """Phenomenological d = 5 x 5 rounds (11-bit frontier): both lowerings route it to the global-memory executor, the
    generic tables are identical and every pass carries a butterfly block."""
    factors, checks, dims = _dem_graph(tq.parse_dem_file(str(golden_dir / "surface_d5_r5_phenom.dem")))
    py = D._lower_sumprod(factors, checks, dims[0], dims[1], dims[2], None)
    lw = _cabi.Lowered(_cabi.Problem(factors, checks, S.SUMPROD, *dims))
    m = lw.meta
    assert m["kind"] == 2 and hasattr(py, "passes") and (m["n_pass"], m["w_cap"]) == (len(py.passes), py.w_cap)
    for what, arr in [(_cabi.LW_WD_PASS_HDR, py.pass_hdr), (_cabi.LW_WD_STEP_HDR, py.step_hdr), (_cabi.LW_WD_INTS, py.ints),
                      (_cabi.LW_WD_TABLES, py.tables)]:
        assert _same(lw.get(what), arr), what
    assert (lw.get(_cabi.LW_WD_BF_OFF) >= 0).all()


@pytest.mark.parametrize("t_max", [8, 12])
def test_wide_tables_identical(t_max, monkeypatch):
    txt = tq.surface_memory_circuit(3, 3, "Z", 2e-3, 2e-3, 2e-3, 2e-3)
    factors, checks, dims = _dem_graph(tq.detector_error_model(tq.parse_stim_string(txt)))
    monkeypatch.setenv("TQEC_FORCE_WIDE", "1")
    monkeypatch.setenv("TQEC_WIDE_TMAX", str(t_max))
    py = D._lower_sumprod(factors, checks, dims[0], dims[1], dims[2], None)
    monkeypatch.delenv("TQEC_FORCE_WIDE")
    monkeypatch.delenv("TQEC_WIDE_TMAX")
    lw = _cabi.Lowered(_cabi.Problem(factors, checks, S.SUMPROD, *dims, flags=_cabi.COMPILE_FORCE_WIDE, wide_t_max=t_max))
    m = lw.meta
    assert m["kind"] == 2 and (m["n_pass"], m["w_cap"], m["t_max"], m["log2_scale"]) == (len(py.passes), py.w_cap, t_max, py.log2_scale)
    assert _same(lw.get(_cabi.LW_ORDER), py.order)
    for what, arr in [(_cabi.LW_WD_PASS_HDR, py.pass_hdr), (_cabi.LW_WD_STEP_HDR, py.step_hdr), (_cabi.LW_WD_INTS, py.ints),
                      (_cabi.LW_WD_TABLES, py.tables), (_cabi.LW_WD_OBS_POS, py.obs_pos)]:
        assert _same(lw.get(what), arr), what
    cost = lw.get(_cabi.LW_COST)
    assert cost[0] == py.cost and cost[1] == py.bytes_per_shot


def test_library_orders_the_d5_circuit_level_dem_itself():
    """BASELINE configs[3], d = 5 x 5 rounds: the library's own ordering (spectral sweep, Jacobi eigen-solver) finds a
    front of at most 29 bits and lowers the 1605 mechanisms to global-memory passes; with the Python order the tables
    are identical."""
    txt = tq.surface_memory_circuit(5, 5, "Z", 1e-3, 1e-3, 1e-3, 1e-3)
    factors, checks, dims = _dem_graph(tq.detector_error_model(tq.parse_stim_string(txt)))
    lw = _cabi.Lowered(_cabi.Problem(factors, checks, S.SUMPROD, *dims))
    m = lw.meta
    assert m["kind"] == 2 and m["w_cap"] <= 29 and m["wide_steps"] == 1605
    order = [int(i) for i in lw.get(_cabi.LW_ORDER)]
    py = D._lower_sumprod(factors, checks, dims[0], dims[1], dims[2], order)
    assert _same(lw.get(_cabi.LW_WD_PASS_HDR), py.pass_hdr) and _same(lw.get(_cabi.LW_WD_TABLES), py.tables)
    assert _same(lw.get(_cabi.LW_WD_INTS), py.ints) and _same(lw.get(_cabi.LW_WD_STEP_HDR), py.step_hdr)


def test_lowering_errors_are_reported():
    import ctypes as C
    f = [S.Factor((0,), np.array([0.9, 0.1]))]
    with pytest.raises(_cabi.TqecError, match="variable id out of range"):
        _cabi.Lowered(_cabi.Problem(f, [S.Check((3,), "syn", 0)], S.MAXPLUS, 1, 1, 0))
    with pytest.raises(_cabi.TqecError, match="finite and non-negative"):
        _cabi.Lowered(_cabi.Problem([S.Factor((0,), np.array([1.5, -0.5]))], [S.Check((0,), "syn", 0)], S.MAXPLUS, 1, 1, 0))
    with pytest.raises(_cabi.TqecError, match="permutation"):
        _cabi.Lowered(_cabi.Problem(f, [S.Check((0,), "syn", 0)], S.MAXPLUS, 1, 1, 0, order=[1]))
    with pytest.raises(_cabi.TqecError, match="observable"):
        _cabi.Lowered(_cabi.Problem(f, [S.Check((0,), "syn", 0)], S.SUMPROD, 1, 1, 1))
    h = C.c_void_p()
    assert _cabi.lib().tqec_lower(None, C.byref(h)) == -1 and not h.value


@pytest.mark.parametrize("case", ["tnmap_d7", "tnmmap_d5", "dem_wide", "steane"])
def test_lowered_plan_round_trips_through_a_file(case, tmp_path):
    """tqec_lowered_save / tqec_lowered_load: every table the plan is created from comes back identical (sweep plan with its
    head tables, sum-product schedule, global-memory passes with butterfly blocks, a small general schedule); files that are
    truncated, corrupt or not plans at all are refused."""
    import os
    from tensorqec.jl_b200._cabi import TqecError
    if case == "tnmap_d7":
        factors, checks, nq, ns = _tnmap_graph(tq.SurfaceCode(7, 7))
        prob = _cabi.Problem(factors, checks, S.MAXPLUS, nq, ns, 0, head_bits=10)
    elif case == "steane":
        factors, checks, nq, ns = _tnmap_graph(tq.SteaneCode())
        prob = _cabi.Problem(factors, checks, S.MAXPLUS, nq, ns, 0)
    elif case == "tnmmap_d5":
        t = tq.CSSTannerGraph(tq.SurfaceCode(5, 5))
        _, _, factors, checks, dims, _, _, _ = D._tnmmap_css_graph(tq.get_problem(t, tq.iid_error(0.05, t)))
        prob = _cabi.Problem(factors, checks, S.SUMPROD, dims[0], dims[1], dims[2])
    else:
        dem = tq.parse_dem_file(os.path.join(os.path.dirname(__file__), "golden", "surface_d3_r3_phenom.dem"))
        factors, checks, dims = _dem_graph(dem)
        prob = _cabi.Problem(factors, checks, S.SUMPROD, dims[0], dims[1], dims[2], flags=_cabi.COMPILE_FORCE_WIDE)
    lw = _cabi.Lowered(prob)
    path = tmp_path / "plan.tqlw"
    lw.save(path)
    back = _cabi.Lowered.load(path)
    assert back.meta == lw.meta
    for what, dt in [(w, _cabi._LW_DTYPE.get(w, np.int32)) for w in range(24)]:
        a, b = lw.get(what), back.get(what)
        assert a.dtype == b.dtype == np.dtype(dt) and _same(a, b), what
    raw = path.read_bytes()
    for name, blob in (("cut", raw[: len(raw) // 2]), ("tail", raw[:-3]), ("magic", b"NOTAPLAN" + raw[8:]),
                       ("format", raw[:8] + (999).to_bytes(4, "little") + raw[12:]), ("empty", b"")):
        bad = tmp_path / f"{name}.tqlw"
        bad.write_bytes(blob)
        with pytest.raises(TqecError):
            _cabi.Lowered.load(bad)
    with pytest.raises(TqecError):
        _cabi.Lowered.load(tmp_path / "missing.tqlw")
