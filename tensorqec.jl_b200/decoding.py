"""The decoder plugin boundary: problems, `TNMAP` / `TNMMAP`, `compile`, `decode`.

Reference: src/decoding/interfaces.jl:6-51 (problem structs), :58-67 (decoder / compiled-decoder abstract types),
:74-79 (`compile` overloads), :96-105 (`decode` overloads), :116-119 (`DecodingResult`), :124-142 (CSS / classical
-> general reduction); src/decoding/general_decoding.jl:5-32 (`single_qubit_tensor`, `reduce2general`,
`extract_decoding`); src/decoding/tndecoder.jl:9-57 (TNMAP), :64-174 (TNMMAP, CSS), :176-271 (TNMMAP, DEM).

`compile` builds the decoder's factor graph exactly as the reference builds its tensor network (same variables, same
prior tensors, same parity constraints, same open logical axes), lowers it to a frontier schedule
(schedule.py) and uploads it through `tqec_plan_create`; `decode` marshals bit-packed syndromes through
`tqec_decode_map` / `tqec_decode_marginal` (+ `tqec_coset_rep`).  `decode` accepts one syndrome or a batch.
"""
from __future__ import annotations

import os

from dataclasses import dataclass, field
from typing import Any, List, Optional, Sequence

import numpy as np

from . import _cabi, schedule as S
from .sweep import lower_sweep
from .wide import lower_wide
from .dem import DetectorErrorModel, dem2tanner
from .error_model import (AbstractErrorModel, CSSErrorPattern, CSSSyndrome, IndependentDepolarizingError,
                          IndependentFlipError, SimpleSyndrome, iid_error)
from .mod2 import ValidatedBits, as_bits, concat_bits, pack_bits, unpack_bits
from .tanner import (AbstractTannerGraph, CSSTannerGraph, SimpleTannerGraph, gf2_right_inverse, gf2_sector_fixes,
                     logical_operator)


# ---- problems (interfaces.jl:6-51) ---------------------------------------------------------------------------
class AbstractDecodingProblem:
    pass


@dataclass
class SimpleTensorNetwork:
    """Container of the prior distribution (src/nonclifford/tensornetwork.jl:143-146): `ixs[t]` = 0-based variable
    labels of tensor t, `tensors[t]` = ndarray with one axis per label (T[i0, i1, ...] = the reference's
    T[i0+1, i1+1, ...])."""
    ixs: List[List[int]]
    tensors: List[np.ndarray]


@dataclass
class ClassicalDecodingProblem(AbstractDecodingProblem):
    tanner: SimpleTannerGraph
    pvec: IndependentFlipError

    def __post_init__(self):
        if not isinstance(self.pvec, IndependentFlipError):
            self.pvec = IndependentFlipError(self.pvec)


@dataclass
class IndependentDepolarizingDecodingProblem(AbstractDecodingProblem):
    tanner: CSSTannerGraph
    pvec: IndependentDepolarizingError


@dataclass
class GeneralDecodingProblem(AbstractDecodingProblem):
    tanner: SimpleTannerGraph
    ptn: SimpleTensorNetwork


def get_problem(tanner, pvec):
    if isinstance(tanner, SimpleTannerGraph) and isinstance(pvec, IndependentFlipError):
        return ClassicalDecodingProblem(tanner, pvec)
    if isinstance(tanner, CSSTannerGraph) and isinstance(pvec, IndependentDepolarizingError):
        return IndependentDepolarizingDecodingProblem(tanner, pvec)
    raise TypeError(f"no decoding problem for ({type(tanner).__name__}, {type(pvec).__name__})")   # MethodError


# ---- decoders ---------------------------------------------------------------------------------------------------
class AbstractDecoder:
    pass


class AbstractClassicalDecoder(AbstractDecoder):
    pass


class AbstractGeneralDecoder(AbstractDecoder):
    pass


class CompiledDecoder:
    pass


@dataclass
class DecodingResult:
    """interfaces.jl:116-119, plus what the batched kernels know anyway: `logp` (TNMAP: log-weight of the returned
    configuration), `marginal` / `sector` (TNMMAP: logical-sector weights and their argmax)."""
    success_tag: Any
    error_pattern: Any
    logp: Optional[np.ndarray] = None
    marginal: Optional[np.ndarray] = None
    sector: Optional[np.ndarray] = None


@dataclass
class TNMAP(AbstractGeneralDecoder):
    """tndecoder.jl:9-11.  `optimizer`: None = built-in frontier ordering; a sequence = absorption order of the
    prior factors (e.g. the leaf order of a contraction tree found by the caller's TreeSA / GreedyMethod)."""
    optimizer: Any = None
    device: int = 0
    table_bits: int = 16         # plans with at most this many syndrome bits are decoded once per syndrome at compile time
                                 # and served from that table (k_lookup); up to 26 (table of 2^bits entries)
    head_bits: int = 14          # syndrome bits the tabulated head of the sweep lowering may depend on (sweep.py): more
                                 # bits = fewer steps per shot, a larger table (2^bits x 2^W entries) and a slower compile

    def __repr__(self):
        return "TNMAP"


@dataclass
class NoOptimizer:
    """tndecoder.jl:64: keep the natural order of the factors."""

    def __repr__(self):
        return "NoOptimizer"


@dataclass
class TNMMAP(AbstractGeneralDecoder):
    """tndecoder.jl:77-80.  `factorize` only affects how the reference shapes its DEM network (:195, :199); the
    frontier lowering always uses the fully factorised (XOR-chain) form, which contracts to the same marginals."""
    optimizer: Any = None
    factorize: bool = True
    device: int = 0
    table_bits: int = 16         # as for TNMAP: problems with at most this many syndrome / detector bits are tabulated
    head_bits: int = 14          # as for TNMAP (CSS plans on k_sweep<SUMPROD>: 2^bits x 2^W doubles, no configuration table)
    dynamic_rescale: bool = False  # per-shot dynamic rescaling (int32 exponent per shot) against FP64 underflow on extremely
                                 # unlikely syndromes; runs the plan on the global-memory executor (slower; tables are static-
                                 # ally scaled in any case, which covers every BASELINE config)

    def __repr__(self):
        return "TNMMAP"


def _order_of(optimizer, n_factors):
    if optimizer is None:
        return None
    if isinstance(optimizer, NoOptimizer):
        return list(range(n_factors))
    return [int(i) for i in optimizer]


# ---- reductions (general_decoding.jl) ----------------------------------------------------------------------------
def single_qubit_tensor(px, py, pz) -> np.ndarray:
    """general_decoding.jl:5: T[x, z] = [1-px-py-pz  pz ; px  py]."""
    return np.array([[1.0 - px - py - pz, pz], [px, py]])


@dataclass
class CSSToGeneralDecodingProblem:
    qubit_num: int


def reduce2general(tanner: CSSTannerGraph, pvec_or_tn):
    """general_decoding.jl:20-32: 2n variables (x_i = i, z_i = i + n); checks = X checks on the z block, then Z checks
    on the x block; priors = one 2x2 tensor per qubit on (i, i+n) or a caller-supplied network."""
    n = tanner.stgx.nq
    if isinstance(pvec_or_tn, IndependentDepolarizingError):
        p = pvec_or_tn
        tn = SimpleTensorNetwork([[i, i + n] for i in range(n)],
                                 [single_qubit_tensor(p.px[j], p.py[j], p.pz[j]) for j in range(n)])
    else:
        tn = pvec_or_tn
    sts = [[q + n for q in s] for s in tanner.stgx.s2q] + [list(s) for s in tanner.stgz.s2q]
    return GeneralDecodingProblem(SimpleTannerGraph(2 * n, sts), tn), CSSToGeneralDecodingProblem(n)


def extract_decoding(cgdp: CSSToGeneralDecodingProblem, error_pattern: np.ndarray) -> DecodingResult:
    n = cgdp.qubit_num
    e = np.asanyarray(error_pattern)                            # keeps the ValidatedBits mark of the library's own corrections
    return DecodingResult(True, CSSErrorPattern(e[..., :n], e[..., n:2 * n]))


# ---- TNMAP (tndecoder.jl:16-57) -------------------------------------------------------------------------------------
def _table_bits_abi(tb) -> int:
    """decoder.table_bits -> ABI value (0 = library default, -1 = never)."""
    return 0 if tb is None else (-1 if int(tb) <= 0 else int(tb))


class _LazySchedule:
    """`compile` lowers the factor graph INSIDE libtqec_cuda.so (tqec_lower, csrc/tqec_lower*.cpp).  The Python
    lowering (schedule.py / sweep.py / wide.py) is the test oracle of that code and emits bit-identical tables
    (tests/test_lower_cpp.py); `.schedule` runs it on demand -- with the absorption order the library chose -- for
    tests and inspection.  TQEC_PY_LOWERING=1 makes `compile` upload the Python tables instead."""

    def _init_plan(self, problem_args, py_lower, device, decoder):
        self._py_lower = py_lower
        self._schedule = None
        if os.environ.get("TQEC_PY_LOWERING") is not None:
            self._schedule = py_lower(None)
            self.plan = _cabi.Plan(self._schedule, device)
        else:
            factors, checks, semiring, n_vars, n_checks, n_obs, order = problem_args
            dyn = bool(getattr(decoder, "dynamic_rescale", False))
            prob = _cabi.Problem(factors, checks, semiring, n_vars, n_checks, n_obs, order=order,
                                 head_bits=int(getattr(decoder, "head_bits", 0) or 0),
                                 table_bits=-1 if dyn else _table_bits_abi(decoder.table_bits), device=device,
                                 flags=_cabi.COMPILE_DYNAMIC_RESCALE if dyn else 0)
            self.plan = _cabi.Plan.compile(prob)
            lw = self.plan.lowered
            if lw["kind"] == 0 and lw["w_max"] >= 5 and not self.plan.query(_cabi.Q_TABLE):
                import warnings
                warnings.warn(f"{type(decoder).__name__}: this plan (frontier {lw['w_max']} bits) does not fit the in-place patch sweep "
                              f"(k_sweep) and decodes through the general frontier kernels, at roughly half the throughput; "
                              f"rotated surface codes (d = 5 .. 9, even and rectangular ones included) in the library's own order do fit", RuntimeWarning, stacklevel=4)

    @property
    def schedule(self):
        if self._schedule is None:
            self._schedule = self._py_lower(self.plan.order)
        return self._schedule


class CompiledTNMAP(CompiledDecoder, _LazySchedule):
    def __init__(self, decoder: "TNMAP", problem: "GeneralDecodingProblem"):
        factors, checks = _tnmap_graph(problem)
        t = problem.tanner
        self.qubit_num = t.nq
        self.n_checks = t.ns
        self.device = decoder.device
        order = _order_of(decoder.optimizer, len(factors))
        self._init_plan((factors, checks, S.MAXPLUS, t.nq, t.ns, 0, order),
                        lambda o: _tnmap_lower(decoder, factors, checks, t.nq, t.ns, order if o is None else o),
                        decoder.device, decoder)


def _tnmap_graph(problem: GeneralDecodingProblem):
    """The factor graph of compile(::TNMAP, ::GeneralDecodingProblem) (tndecoder.jl:33-50): one factor per prior tensor,
    one clamped parity row per check."""
    t = problem.tanner
    factors = [S.Factor(tuple(int(v) for v in ix), S.flat_table(tt)) for ix, tt in zip(problem.ptn.ixs, problem.ptn.tensors)]
    for f in factors:
        if any(not 0 <= v < t.nq for v in f.vars):
            raise IndexError("prior tensor label outside 0..nq-1")
    checks = [S.Check(tuple(c), "syn", s) for s, c in enumerate(t.s2q)]
    return factors, checks


def _tnmap_lower(decoder, factors, checks, nq, ns, order) -> S.Schedule:
    """Python lowering of a TNMAP factor graph (the oracle of tqec_lower for max-plus plans)."""
    if os.environ.get("TQEC_NO_SWEEP") is None:
        # preferred lowering: the in-place patch sweep of the unfused schedule (sweep.py); plans whose steps do not fit
        # its shapes fall through to the general kernels
        try:
            su = S.lower(factors, checks, S.MAXPLUS, nq, ns, 0, order=order, fuse=False)
            sw = lower_sweep(su, max_head_bits=int(os.environ.get("TQEC_HEAD_BITS", decoder.head_bits))) if int(os.environ.get("TQEC_SWEEP_MINW", "5")) <= su.w_max <= 10 else None
        except ValueError:
            sw = None
        if sw is not None:
            su.sweep = sw
            su.table_bits = decoder.table_bits
            return su
    sch = S.lower(factors, checks, S.MAXPLUS, nq, ns, 0, order=order)
    sch.table_bits = decoder.table_bits
    return sch


def tnmap_schedule(decoder: TNMAP, problem: GeneralDecodingProblem) -> S.Schedule:
    """Host-only part of compile(::TNMAP, ::GeneralDecodingProblem) in Python: factor graph -> lowered schedule."""
    factors, checks = _tnmap_graph(problem)
    return _tnmap_lower(decoder, factors, checks, problem.tanner.nq, problem.tanner.ns, _order_of(decoder.optimizer, len(factors)))


def _lower_sumprod(factors, checks, n_vars, n_checks, n_obs, order, dynamic=False):
    """Sum-product plan of a marginal network: the on-chip schedule (schedule.py) when the frontier has at most 11 bits
    (9 for plans of rank-1 factors), else the global-memory lowering (wide.py; measured faster from there on: a plan of
    up to 12 bits is a single tile).  The order is chosen once and shared by both."""
    all_check_vars = {v for c in checks for v in c.vars}
    merged = S.merge_overlapping(list(factors), n_vars, all_check_vars, allow_negative=True)
    if order is None:
        order = S.choose_order(merged, checks)
    elif len(order) == len(factors) and len(factors) != len(merged):
        order = S.map_order(factors, merged, order)
    w_max, _ = S._evaluate(order, S._Sim(merged, checks))
    # plans made of rank-1 factors only (detector error models) run as register butterflies on the global-memory
    # executor (k_wide_bf, library lowering): measured faster from 10 bits on
    dflt = 9 if all(len(f.vars) == 1 for f in merged) else S.MAX_SUMPROD_ONCHIP_WIDTH
    onchip = min(int(os.environ.get("TQEC_SUMPROD_ONCHIP_WIDTH", dflt)), S.MAX_SMEM_WIDTH)
    if w_max <= onchip and os.environ.get("TQEC_FORCE_WIDE") is None and not dynamic:
        return S.lower(merged, checks, S.SUMPROD, n_vars, n_checks, n_obs, order=order)
    return lower_wide(merged, checks, S.SUMPROD, n_vars, n_checks, n_obs, order=order,
                      t_max=int(os.environ.get("TQEC_WIDE_TMAX", "12")), max_drop_bits=600.0 if dynamic else 0.0)


def _attach_sweep(sch, max_head_bits: int = 10):
    """Sum-product plans are never fused, so the schedule itself is what `lower_sweep` expects; plans it declines keep
    running on the general kernels."""
    if os.environ.get("TQEC_NO_SWEEP") is None and hasattr(sch, "steps") and 5 <= sch.w_max <= 10:
        try:
            sw = lower_sweep(sch, max_head_bits=max_head_bits)
        except ValueError:
            sw = None
        if sw is not None:
            sch.sweep = sw


def _compile_tnmap(decoder: TNMAP, problem: GeneralDecodingProblem) -> CompiledTNMAP:
    return CompiledTNMAP(decoder, problem)


def _syndrome_bits(s, n) -> np.ndarray:
    b = s if isinstance(s, ValidatedBits) else as_bits(s)
    b = b[None, :] if b.ndim == 1 else b
    if b.shape[1] != n:
        raise ValueError(f"syndrome has {b.shape[1]} bits, the decoder expects {n}")
    return b


def _decode_tnmap(ct: CompiledTNMAP, syndrome: SimpleSyndrome) -> DecodingResult:
    raw = syndrome.s
    bits = raw if isinstance(raw, ValidatedBits) else as_bits(raw).view(ValidatedBits)   # validated once: 1e7 shots are 0.8 GB
    single = bits.ndim == 1
    bits = _syndrome_bits(bits, ct.n_checks)
    cfg, logp = ct.plan.decode_map_bits(bits, ct.qubit_num)     # one byte per bit both ways; packed on the device
    cfg = cfg.view(ValidatedBits)                                # the library's own 0/1 bytes: no validation pass downstream
    ok = np.isfinite(logp)
    if single:
        return DecodingResult(bool(ok[0]), cfg[0], logp=logp[0])
    return DecodingResult(ok, cfg, logp=logp)


@dataclass
class CompiledGeneralDecoder(CompiledDecoder):
    """interfaces.jl:124-127."""
    cd: CompiledDecoder
    reduction: CSSToGeneralDecodingProblem


# ---- TNMMAP, CSS (tndecoder.jl:85-174) --------------------------------------------------------------------------------
class CompiledTNMMAP(CompiledDecoder, _LazySchedule):
    def __init__(self, decoder: "TNMMAP", problem: "IndependentDepolarizingDecodingProblem"):
        lx, lz, factors, checks, dims, R, L, FIX = _tnmmap_css_graph(problem)
        self.tanner, self.lx, self.lz = problem.tanner, lx, lz
        n_vars, n_checks, self.n_obs = dims
        device = decoder.device
        order = _order_of(decoder.optimizer, len(factors))
        self._init_plan((factors, checks, S.SUMPROD, n_vars, n_checks, self.n_obs, order),
                        lambda o: _sumprod_lower(decoder, factors, checks, dims, order if o is None else o), device, decoder)
        self.R = _cabi.GF2Matrix(R, device)
        self.L = _cabi.GF2Matrix(L, device)
        self.FIX = _cabi.GF2Matrix(FIX, device)
        self.device = device

    def marginal(self, syndrome: CSSSyndrome) -> np.ndarray:
        """`ct.code(ct.tensors...)` after `update_syndrome!` (tndecoder.jl:148-162): array with 2k axes of size 2,
        first k = lx-parities of the Z errors, last k = lz-parities of the X errors (batch axis first if batched)."""
        single = as_bits(syndrome.sx).ndim == 1
        bits = np.concatenate([_syndrome_bits(syndrome.sx, self.tanner.stgx.ns),
                               _syndrome_bits(syndrome.sz, self.tanner.stgz.ns)], axis=1)
        mar, _ = self.plan.decode_marginal(pack_bits(bits))
        k2 = self.n_obs
        out = mar.reshape((mar.shape[0],) + (2,) * k2, order="F") if k2 else mar
        return out[0] if single else out


def _sumprod_lower(decoder, factors, checks, dims, order):
    """Python lowering of a marginal network (the oracle of tqec_lower for sum-product plans)."""
    sch = _lower_sumprod(factors, checks, dims[0], dims[1], dims[2], order, dynamic=bool(getattr(decoder, "dynamic_rescale", False)))
    _attach_sweep(sch, max_head_bits=int(os.environ.get("TQEC_HEAD_BITS_SP", getattr(decoder, "head_bits", 0) or 14)))
    sch.table_bits = decoder.table_bits
    return sch


def _tnmmap_css_graph(problem: IndependentDepolarizingDecodingProblem):
    """The marginal network of compile(::TNMMAP, CSS) (tndecoder.jl:97-146) as a factor graph, and the matrices of
    error_pattern (:167-174) -> (lx, lz, factors, checks, (n_vars, n_checks, n_obs), R, L, FIX)."""
    tanner = problem.tanner
    n = tanner.stgx.nq
    nsx, nsz = tanner.stgx.ns, tanner.stgz.ns
    lx, lz = logical_operator(tanner)
    k = lx.shape[0]
    p = problem.pvec
    factors = [S.Factor((i, i + n), S.flat_table(single_qubit_tensor(p.px[i], p.py[i], p.pz[i]))) for i in range(n)]
    checks = [S.Check(tuple(q + n for q in c), "syn", i) for i, c in enumerate(tanner.stgx.s2q)]
    checks += [S.Check(tuple(c), "syn", nsx + i) for i, c in enumerate(tanner.stgz.s2q)]
    # open axes, in the reference's output order iy (tndecoder.jl:134): lx-parities of Z errors, then lz-parities of X errors
    checks += [S.Check(tuple(int(q) + n for q in np.flatnonzero(lx[i])), "obs", i) for i in range(k)]
    checks += [S.Check(tuple(int(q) for q in np.flatnonzero(lz[i])), "obs", k + i) for i in range(k)]
    # error_pattern (tndecoder.jl:167-174): any solution of the syndrome equations, moved into the decoded sector
    Rz, _ = gf2_right_inverse(tanner.stgz.H)            # ex = Rz sz
    Rx, _ = gf2_right_inverse(tanner.stgx.H)            # ez = Rx sx
    R = np.zeros((2 * n, nsx + nsz), dtype=np.uint8)
    R[:n, nsx:] = Rz
    R[n:, :nsx] = Rx
    L = np.zeros((2 * k, 2 * n), dtype=np.uint8)
    FIX = np.zeros((2 * k, 2 * n), dtype=np.uint8)
    L[:k, n:] = lx                                       # sector bit i     = lx[i] . ez ; repaired by ez += lz[i]
    FIX[:k, n:] = lz
    L[k:, :n] = lz                                       # sector bit k + i = lz[i] . ex ; repaired by ex += lx[i]
    FIX[k:, :n] = lx
    return lx, lz, factors, checks, (2 * n, nsx + nsz, 2 * k), R, L, FIX


def tnmmap_css_schedule(decoder: TNMMAP, problem: IndependentDepolarizingDecodingProblem):
    """Host-only part of the CSS TNMMAP compile in Python -> (lx, lz, schedule, R, L, FIX)."""
    lx, lz, factors, checks, dims, R, L, FIX = _tnmmap_css_graph(problem)
    return lx, lz, _sumprod_lower(decoder, factors, checks, dims, _order_of(decoder.optimizer, len(factors))), R, L, FIX


def _compile_tnmmap_css(decoder: TNMMAP, problem: IndependentDepolarizingDecodingProblem) -> CompiledTNMMAP:
    return CompiledTNMMAP(decoder, problem)


def _decode_tnmmap_css(ct: CompiledTNMMAP, syndrome: CSSSyndrome) -> DecodingResult:
    single = as_bits(syndrome.sx).ndim == 1
    bits = np.concatenate([_syndrome_bits(syndrome.sx, ct.tanner.stgx.ns), _syndrome_bits(syndrome.sz, ct.tanner.stgz.ns)], axis=1)
    words = pack_bits(bits)
    mar, pos = ct.plan.decode_marginal(words)
    n = ct.tanner.stgx.nq
    ew, in_sector = _cabi.coset_rep(ct.R, ct.L, ct.FIX, words, pos)
    e = unpack_bits(ew, 2 * n)
    ok = (mar.max(axis=1) > 0) & in_sector
    k2 = ct.n_obs
    marr = mar.reshape((mar.shape[0],) + (2,) * k2, order="F")
    if single:
        return DecodingResult(bool(ok[0]), CSSErrorPattern(e[0, :n], e[0, n:]), marginal=marr[0], sector=int(pos[0]))
    return DecodingResult(ok, CSSErrorPattern(e[:, :n], e[:, n:]), marginal=marr, sector=pos)


# ---- TNMMAP, detector error model (tndecoder.jl:176-271) -----------------------------------------------------------------
class CompiledDEMTNMMAP(CompiledDecoder, _LazySchedule):
    def __init__(self, decoder: "TNMMAP", dem: DetectorErrorModel):
        tanner, l2q, factors, checks, dims, R, L, FIX = _tnmmap_dem_graph(dem)
        self.tanner, self.l2q = tanner, l2q
        self.n_obs = dims[2]
        device = decoder.device
        order = _order_of(decoder.optimizer, len(factors))
        self._init_plan((factors, checks, S.SUMPROD, dims[0], dims[1], dims[2], order),
                        lambda o: _sumprod_lower(decoder, factors, checks, dims, order if o is None else o), device, decoder)
        self.R = _cabi.GF2Matrix(R, device)
        self.L = _cabi.GF2Matrix(L, device)
        self.FIX = _cabi.GF2Matrix(FIX, device)
        self.device = device

    def marginal(self, syndrome: SimpleSyndrome) -> np.ndarray:
        single = as_bits(syndrome.s).ndim == 1
        mar, _ = self.plan.decode_marginal(pack_bits(_syndrome_bits(syndrome.s, self.tanner.ns)))
        k = self.n_obs
        out = mar.reshape((mar.shape[0],) + (2,) * k, order="F") if k else mar
        return out[0] if single else out


def _tnmmap_dem_graph(dem: DetectorErrorModel):
    """The marginal network of compile(::TNMMAP, ::DetectorErrorModel) (tndecoder.jl:186-219) as a factor graph
    -> (tanner, l2q, factors, checks, (n_vars, n_checks, n_obs), R, L, FIX)."""
    tanner = dem2tanner(dem)
    ne, nd = tanner.nq, tanner.ns
    l2q = [[e for e in range(ne) if l in dem.flipped_detectors[e]] for l in dem.logical_list]
    factors = [S.Factor((e,), np.array([1.0 - p, p])) for e, p in enumerate(dem.error_rates)]
    checks = [S.Check(tuple(c), "syn", d) for d, c in enumerate(tanner.s2q)]
    checks += [S.Check(tuple(c), "obs", l) for l, c in enumerate(l2q)]
    R, _ = gf2_right_inverse(tanner.H)
    L = np.zeros((len(l2q), ne), dtype=np.uint8)
    for l, c in enumerate(l2q):
        L[l, c] = 1
    # The reference "repairs" a wrong observable by flipping every mechanism that touches it (tndecoder.jl:257-258,
    # 266-267), which also changes the detectors.  Here the repair is an undetectable combination of
    # mechanisms (H f = 0) with the required sector flip L f; `gf2_sector_fixes` returns a basis of all reachable flips,
    # so joint flips of several observables are found too, and an unreachable sector is reported (success_tag False)
    # instead of returning a pattern in the wrong sector.
    FIX = gf2_sector_fixes(tanner.H, L)
    return tanner, l2q, factors, checks, (ne, nd, len(l2q)), R, L, FIX


def tnmmap_dem_schedule(decoder: TNMMAP, dem: DetectorErrorModel):
    """Host-only part of the DEM TNMMAP compile in Python -> (tanner, l2q, schedule, R, L, FIX)."""
    tanner, l2q, factors, checks, dims, R, L, FIX = _tnmmap_dem_graph(dem)
    return tanner, l2q, _sumprod_lower(decoder, factors, checks, dims, _order_of(decoder.optimizer, len(factors))), R, L, FIX


def _compile_tnmmap_dem(decoder: TNMMAP, dem: DetectorErrorModel) -> CompiledDEMTNMMAP:
    return CompiledDEMTNMMAP(decoder, dem)


def _decode_tnmmap_dem(ct: CompiledDEMTNMMAP, syndrome: SimpleSyndrome) -> DecodingResult:
    single = as_bits(syndrome.s).ndim == 1
    words = pack_bits(_syndrome_bits(syndrome.s, ct.tanner.ns))
    mar, pos = ct.plan.decode_marginal(words)
    ew, in_sector = _cabi.coset_rep(ct.R, ct.L, ct.FIX, words, pos)
    e = unpack_bits(ew, ct.tanner.nq)
    ok = (mar.max(axis=1) > 0) & in_sector                      # False: no undetectable pattern reaches the decoded sector
    k = ct.n_obs
    marr = mar.reshape((mar.shape[0],) + (2,) * k, order="F")
    if single:
        return DecodingResult(bool(ok[0]), e[0], marginal=marr[0], sector=int(pos[0]))
    return DecodingResult(ok, e, marginal=marr, sector=pos)


# ---- compile / decode dispatch (interfaces.jl:74-105, 129-142) ----------------------------------------------------------
def compile(decoder: AbstractDecoder, problem, pvec: Optional[AbstractErrorModel] = None):
    """compile(decoder, problem) | compile(decoder, tanner) | compile(decoder, tanner, pvec) | compile(TNMMAP, dem)."""
    if isinstance(problem, AbstractTannerGraph):
        if pvec is None:
            pvec = iid_error(0.05, problem)                                    # interfaces.jl:74-76
        problem = get_problem(problem, pvec)                                   # interfaces.jl:77-79
    elif pvec is not None:
        raise TypeError("pvec is only meaningful together with a Tanner graph")
    if isinstance(decoder, TNMAP):
        if isinstance(problem, GeneralDecodingProblem):
            return _compile_tnmap(decoder, problem)
        if isinstance(problem, IndependentDepolarizingDecodingProblem):        # interfaces.jl:129-133
            gdp, c2g = reduce2general(problem.tanner, problem.pvec)
            return CompiledGeneralDecoder(_compile_tnmap(decoder, gdp), c2g)
        if isinstance(problem, ClassicalDecodingProblem):                      # interfaces.jl:139-142
            t = problem.tanner
            tn = SimpleTensorNetwork([[i] for i in range(t.nq)], [np.array([1.0 - p, p]) for p in problem.pvec.p])
            return _compile_tnmap(decoder, GeneralDecodingProblem(t, tn))
    from .bposd import BPDecoder, compile_bp
    if isinstance(decoder, BPDecoder) and isinstance(problem, ClassicalDecodingProblem):
        return compile_bp(decoder, problem)                                    # bposd.jl:30-35
    from .truthtable import TableDecoder, compile_table
    if isinstance(decoder, TableDecoder) and isinstance(problem, IndependentDepolarizingDecodingProblem):
        return compile_table(decoder, problem)                                 # truthtable.jl:205-208
    if isinstance(decoder, TNMMAP):
        if isinstance(problem, IndependentDepolarizingDecodingProblem):
            return _compile_tnmmap_css(decoder, problem)
        if isinstance(problem, DetectorErrorModel):
            return _compile_tnmmap_dem(decoder, problem)
    raise TypeError(f"no method compile({type(decoder).__name__}, {type(problem).__name__})")


def decode(first, *args):
    """decode(compiled, syndrome) | decode(decoder, problem|tanner, syndrome[, pvec]) (interfaces.jl:96-105)."""
    if isinstance(first, AbstractDecoder):
        if len(args) == 2:
            prob, syn = args
            ct = compile(first, prob)
        elif len(args) == 3:
            tanner, syn, pvec = args
            ct = compile(first, tanner, pvec)
        else:
            raise TypeError("decode(decoder, problem|tanner, syndrome[, pvec])")
        return decode(ct, syn)
    (syn,) = args
    ct = first
    if isinstance(ct, CompiledGeneralDecoder):                                  # interfaces.jl:135-137
        if not isinstance(syn, CSSSyndrome):
            raise TypeError("a CSS-compiled decoder decodes a CSSSyndrome")
        sx, sz = as_bits(syn.sx), as_bits(syn.sz)
        if isinstance(ct.cd, CompiledTNMAP) and sx.ndim == 2 and sz.ndim == 2 and sx.shape[0] == sz.shape[0]:
            # batched TNMAP: the library reads sx and sz where they lie (tqec_decode_map_bytes2), no concatenation here
            if sx.shape[1] + sz.shape[1] != ct.cd.n_checks:
                raise ValueError(f"syndrome has {sx.shape[1] + sz.shape[1]} bits, the decoder expects {ct.cd.n_checks}")
            cfg, logp = ct.cd.plan.decode_map_bits2(sx, sz, ct.cd.qubit_num)
            res = DecodingResult(np.isfinite(logp), cfg.view(ValidatedBits), logp=logp)
        else:
            res = decode(ct.cd, SimpleSyndrome(concat_bits(sx, sz).view(ValidatedBits)))
        out = extract_decoding(ct.reduction, res.error_pattern)
        out.success_tag, out.logp = res.success_tag, res.logp
        return out
    from .bposd import CompiledBP, decode_bp
    if isinstance(ct, CompiledBP) and isinstance(syn, SimpleSyndrome):
        return decode_bp(ct, syn)
    from .truthtable import CompiledTable, decode_table
    if isinstance(ct, CompiledTable) and isinstance(syn, CSSSyndrome):
        return decode_table(ct, syn)
    if isinstance(ct, CompiledTNMAP) and isinstance(syn, SimpleSyndrome):
        return _decode_tnmap(ct, syn)
    if isinstance(ct, CompiledTNMMAP) and isinstance(syn, CSSSyndrome):
        return _decode_tnmmap_css(ct, syn)
    if isinstance(ct, CompiledDEMTNMMAP) and isinstance(syn, SimpleSyndrome):
        return _decode_tnmmap_dem(ct, syn)
    raise TypeError(f"no method decode({type(ct).__name__}, {type(syn).__name__})")
