"""In-place patch sweep (tensorqec.jl_b200/sweep.py) on CPU: the lowering executed by the table-level emulator against
the independent recurrence oracle, the tabulated head, the plan geometry, and the menu of register shapes compiled into
k_sweep (csrc/tqec_sweep_menu.h) against the planner's copy."""
import os
import re

import numpy as np
import pytest

from oracle import frontier, gf2, philox, sweep_emulator

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _syndromes(t, em, seed, B):
    ex, ez = philox.sample_depolarizing(em.px, em.py, em.pz, seed, 0, B)
    sx, sz = gf2.css_syndrome(ex, ez, t.stgx.H, t.stgz.H)
    return np.concatenate([sx, sz], axis=1)


def _plan(tq, d, em=None):
    from tensorqec.jl_b200 import schedule as S, sweep as SW
    t = tq.CSSTannerGraph(tq.SurfaceCode(d, d))
    em = em or tq.iid_error(0.05, t)
    gdp, _ = tq.reduce2general(t, em)
    factors = [S.Factor(tuple(int(v) for v in ix), S.flat_table(tt)) for ix, tt in zip(gdp.ptn.ixs, gdp.ptn.tensors)]
    checks = [S.Check(tuple(c), "syn", s) for s, c in enumerate(gdp.tanner.s2q)]
    su = S.lower(factors, checks, S.MAXPLUS, gdp.tanner.nq, gdp.tanner.ns, 0, fuse=False)
    return t, em, su, SW.lower_sweep(su)


@pytest.mark.parametrize("d,B", [(5, 96), (7, 40), (9, 12)])
def test_sweep_emulator_matches_recurrence_bit_exact(tq, d, B):
    from tensorqec.jl_b200 import sweep as SW
    t, em, su, pl = _plan(tq, d)
    assert pl is not None
    syn = _syndromes(t, em, 100 + d, B)
    syn[0] = 0                                                   # the trivial syndrome
    lp, cfg = sweep_emulator.run(pl, SW.MENU, syn)
    lp0, cfg0 = frontier.run(su.factors, su.checks, su.order, 0, syn, su.n_vars)
    assert np.array_equal(lp, lp0)                               # same IEEE adds in the same order
    assert np.array_equal(cfg, cfg0)                             # same tie rule (smallest assignment)
    H = np.zeros((su.n_checks, su.n_vars), dtype=np.uint8)
    for c in su.checks:
        H[c.index, list(c.vars)] = 1
    assert np.array_equal((cfg @ H.T) % 2, syn)                  # every correction reproduces its syndrome


@pytest.mark.parametrize("dx,dz,B", [(6, 6, 40), (8, 8, 16), (5, 7, 40)])
def test_sweep_fresh_pins_even_and_rectangular_codes(tq, dx, dz, B):
    """Even-distance and rectangular rotated surface codes: their sweeps contain steps that open a check without closing
    one through the same variable (column turns).  Such a FRESH pin takes a dead slot -- a pinned variable whose flip
    mask contains its own bit, so output bit 1 reads the live half -- and a check may close without a successor.  The
    emulator stays bit-identical to the recurrence."""
    from tensorqec.jl_b200 import schedule as S, sweep as SW
    t = tq.CSSTannerGraph(tq.SurfaceCode(dx, dz))
    em = tq.iid_error(0.05, t)
    gdp, _ = tq.reduce2general(t, em)
    factors = [S.Factor(tuple(int(v) for v in ix), S.flat_table(tt)) for ix, tt in zip(gdp.ptn.ixs, gdp.ptn.tensors)]
    checks = [S.Check(tuple(c), "syn", s) for s, c in enumerate(gdp.tanner.s2q)]
    su = S.lower(factors, checks, S.MAXPLUS, gdp.tanner.nq, gdp.tanner.ns, 0, fuse=False)
    pl = SW.lower_sweep(su, 12)
    assert pl is not None, "the plan should fit the in-place patch sweep"
    if dx == dz:
        assert any(any(l.pk and any((m >> pb) & 1 for (_, pb), m in zip(l.pinned, l.pk)) for l in ss.layers) for ss in pl.ssteps), "no fresh pin"
    syn = _syndromes(t, em, 200 + dx, B)
    syn[0] = 0
    lp, cfg = sweep_emulator.run(pl, SW.MENU, syn)
    lp0, cfg0 = frontier.run(su.factors, su.checks, su.order, 0, syn, su.n_vars)
    assert np.array_equal(lp, lp0) and np.array_equal(cfg, cfg0)


@pytest.mark.parametrize("bits", [6, 8, 12])
def test_sweep_head_bits_do_not_change_results(tq, bits):
    """A longer tabulated head (TNMAP(head_bits=...)) removes steps from the per-shot path, not from the arithmetic: the
    emulator's results stay bit-identical to the recurrence for every head length."""
    from tensorqec.jl_b200 import sweep as SW
    t, em, su, _ = _plan(tq, 7)
    pl = SW.lower_sweep(su, max_head_bits=bits)
    assert pl is not None and len(pl.head_bits) <= bits and pl.head_state.shape[0] == 1 << len(pl.head_bits)
    syn = _syndromes(t, em, 31, 24)
    lp, cfg = sweep_emulator.run(pl, SW.MENU, syn)
    lp0, cfg0 = frontier.run(su.factors, su.checks, su.order, 0, syn, su.n_vars)
    assert np.array_equal(lp, lp0) and np.array_equal(cfg, cfg0)


def test_sweep_generic_noise_and_single_shot(tq):
    """Per-qubit noise (no ties): one shot, odd batch sizes, shots that do not fill a team pass."""
    from tensorqec.jl_b200 import sweep as SW
    rng = np.random.default_rng(3)
    t0 = tq.CSSTannerGraph(tq.SurfaceCode(7, 7))
    em = tq.IndependentDepolarizingError(rng.uniform(0.01, 0.1, 49), rng.uniform(0.01, 0.1, 49), rng.uniform(0.01, 0.1, 49))
    t, em, su, pl = _plan(tq, 7, em)
    for B in (1, 9):
        syn = _syndromes(t, em, 5, B)
        lp, cfg = sweep_emulator.run(pl, SW.MENU, syn)
        lp0, cfg0 = frontier.run(su.factors, su.checks, su.order, 0, syn, su.n_vars)
        assert np.array_equal(lp, lp0) and np.array_equal(cfg, cfg0)


@pytest.mark.parametrize("d,B", [(5, 48), (7, 12), (9, 3)])
def test_sweep_sum_product_tnmmap_bit_exact(tq, d, B):
    """TNMMAP (CSS) through the sweep: open observable slots, pinned variables that also flip an observable, a 10-bit
    head table; marginals equal the recurrence oracle bit for bit (same multiplications and additions in the same order)."""
    from tensorqec.jl_b200 import sweep as SW
    t = tq.CSSTannerGraph(tq.SurfaceCode(d, d))
    em = tq.iid_error(0.05, t)
    lx, lz, sch, R, L, FIX = tq.tnmmap_css_schedule(tq.TNMMAP(), tq.get_problem(t, em))
    pl = getattr(sch, "sweep", None)
    assert pl is not None and pl.semiring == 1 and len(pl.out_index) == 4
    syn = _syndromes(t, em, 7 * d, B)
    mar = sweep_emulator.run(pl, SW.MENU, syn)
    ref = frontier.run(sch.factors, sch.checks, sch.order, 1, syn, sch.n_vars)
    assert np.array_equal(mar, np.ldexp(ref, -sch.log2_scale))


@pytest.mark.parametrize("dx,dz,B", [(6, 6, 24), (5, 7, 40), (6, 8, 12)])
def test_sweep_sum_product_even_and_rectangular_codes(tq, dx, dz, B):
    """TNMMAP (CSS) of even-distance / rectangular codes through the sweep (fresh pins + open observable slots)."""
    from tensorqec.jl_b200 import sweep as SW
    t = tq.CSSTannerGraph(tq.SurfaceCode(dx, dz))
    em = tq.iid_error(0.05, t)
    lx, lz, sch, R, L, FIX = tq.tnmmap_css_schedule(tq.TNMMAP(), tq.get_problem(t, em))
    pl = getattr(sch, "sweep", None)
    assert pl is not None and pl.semiring == 1 and len(pl.out_index) == 4
    syn = _syndromes(t, em, 9 * dx + dz, B)
    mar = sweep_emulator.run(pl, SW.MENU, syn)
    ref = frontier.run(sch.factors, sch.checks, sch.order, 1, syn, sch.n_vars)
    assert np.array_equal(mar, np.ldexp(ref, -sch.log2_scale))


def test_sweep_plan_geometry(tq):
    from tensorqec.jl_b200 import sweep as SW
    t, em, su, pl = _plan(tq, 9)
    assert pl.W == 9 and pl.sg == 1 and pl.W + pl.sg == SW.NB
    assert len(pl.head_bits) <= SW.MAX_HEAD_BITS and pl.head_state.shape == (1 << len(pl.head_bits), 1 << pl.W)
    assert pl.head_steps + sum(len(s.layers) for s in pl.ssteps) == len(su.steps)      # every factor absorbed once
    assert pl.conflicts == 0 and not any(s.conflict for s in pl.ssteps)               # bank-conflict-free layout
    for s in pl.ssteps:
        assert len(set(s.pos) | set(s.lanepos) | set(s.looppos)) == len(s.pos) + 5 + len(s.looppos)
        assert len(s.lanepos) == 5 and 0 <= s.menu < len(SW.MENU)
        # lanes cover the four bank-pair bits: positions with distinct residues mod 4 below 8
        assert sorted(p % 4 for p in s.lanepos[:4]) == [0, 1, 2, 3] and all(p < 8 for p in s.lanepos[:4])
    # head table: the all-zero head pattern has a finite best entry; infeasible entries are -inf, never NaN
    assert np.isfinite(pl.head_state[0].max()) and not np.isnan(pl.head_state).any()
    # plans the in-place form cannot express are declined, not mis-lowered
    t3, em3, su3, pl3 = _plan(tq, 3)
    assert pl3 is None


def test_sweep_menu_header_matches_planner():
    from tensorqec.jl_b200 import sweep as SW
    txt = open(os.path.join(ROOT, "tensorqec.jl_b200", "csrc", "tqec_sweep_menu.h")).read()
    rows = re.findall(r"X\(([-\d,\s]+)\)", txt)
    rows = [r for r in rows if len(r.split(",")) == 19]
    assert len(rows) == len(SW.MENU) == int(re.search(r"TQEC_SWEEP_MENU_SIZE (\d+)", txt).group(1))
    for r, (M, layers) in zip(rows, SW.MENU):
        v = [int(x) for x in r.split(",")]
        assert v[0] == rows.index(r) and v[1] == M and v[2] == len(layers)
        for li, layer in enumerate(layers):
            pb, fm = layer[0], layer[1]
            pk = list(layer[2]) if len(layer) > 2 else []
            o = 3 + 8 * li
            assert v[o] == len(pb) and v[o + 3] == len(fm)
            assert [x for x in v[o + 1:o + 3] if x >= 0] == list(pb)
            assert [x for x in v[o + 4:o + 6] if x > 0] == list(fm)
            assert v[o + 6:o + 8] == (pk + [0, 0])[:2]


def test_tnmap_compile_prefers_sweep_and_env_disables_it(tq, monkeypatch):
    t = tq.CSSTannerGraph(tq.SurfaceCode(9, 9))
    gdp, _ = tq.reduce2general(t, tq.iid_error(0.05, t))
    sch = tq.tnmap_schedule(tq.TNMAP(), gdp)
    assert getattr(sch, "sweep", None) is not None and len(sch.steps) == 81
    monkeypatch.setenv("TQEC_NO_SWEEP", "1")
    sch2 = tq.tnmap_schedule(tq.TNMAP(), gdp)
    assert getattr(sch2, "sweep", None) is None
