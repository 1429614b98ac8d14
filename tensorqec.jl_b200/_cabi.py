"""ctypes binding of `libtqec_cuda.so` (C ABI in include/tqec.h).

This is the same binding a Julia maintainer writes with `ccall` (see INTEGRATION.md and julia/TensorQECCUDA.jl).
There is NO fallback: if the library is missing, or no CUDA device is visible when a compute entry point is
called, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TQEC_CUDA_LIB", os.path.join(_HERE, "libtqec_cuda.so"))

OK = 0
MODEL_FLIP, MODEL_DEPOL = 0, 1
(Q_TEAM_THREADS, Q_SHOTS_PER_TEAM, Q_SMEM_BYTES, Q_GRID, Q_TEAMS_PER_SM, Q_BP_BYTES_PER_TEAM, Q_CANDIDATES_PER_SHOT,
 Q_SM_COUNT, Q_LAUNCHES, Q_SWEEP, Q_TABLE, Q_WIDE, Q_WIDE_BATCH) = range(13)

EXPORTS = [
    "tqec_last_error", "tqec_version", "tqec_device_count",
    "tqec_plan_create", "tqec_plan_destroy", "tqec_plan_query",
    "tqec_decode_map", "tqec_decode_map_dev", "tqec_decode_marginal", "tqec_decode_marginal_dev",
    "tqec_gf2_create", "tqec_gf2_destroy", "tqec_gf2_apply", "tqec_gf2_apply_dev",
    "tqec_logical_flags", "tqec_coset_rep", "tqec_sample_errors", "tqec_mc_run", "tqec_fp64_peak",
    "tqec_lower", "tqec_lowered_destroy", "tqec_lowered_get", "tqec_plan_from_lowered", "tqec_plan_compile",
    "tqec_lowered_save", "tqec_lowered_load",
    "tqec_comm_unique_id", "tqec_comm_init", "tqec_comm_destroy", "tqec_comm_allreduce_counts",
    "tqec_decode_map_bytes", "tqec_decode_map_bytes2", "tqec_decode_marginal_bytes", "tqec_dmma_peak", "tqec_decode_marginal_log2",
    "tqec_table_create", "tqec_table_destroy", "tqec_table_decode", "tqec_bp_create", "tqec_bp_destroy", "tqec_bp_decode",
]


class TqecError(RuntimeError):
    pass


class SweepDesc(C.Structure):
    _fields_ = [("W", C.c_int32), ("sg", C.c_int32), ("n_ss", C.c_int32), ("n_head_bits", C.c_int32),
                ("bp_words", C.c_int32), ("n_tvals", C.c_int32),
                ("rec", C.c_void_p), ("tb", C.c_void_p), ("lanetab", C.c_void_p), ("tvals", C.c_void_p),
                ("head_bits", C.c_void_p), ("head_state", C.c_void_p), ("head_cfg", C.c_void_p),
                ("out_index", C.c_void_p)]


class WideDesc(C.Structure):
    _fields_ = [("n_pass", C.c_int32), ("n_steps", C.c_int32), ("w_cap", C.c_int32), ("t_max", C.c_int32),
                ("pass_hdr", C.c_void_p), ("step_hdr", C.c_void_p), ("ints", C.c_void_p), ("n_ints", C.c_int64),
                ("tables", C.c_void_p), ("n_tables", C.c_int64), ("obs_pos", C.c_void_p),
                # butterfly encoding of rank-1 passes: produced by the library's own lowering only (NULL from Python)
                ("bf_off", C.c_void_p), ("bf_ints", C.c_void_p), ("n_bf_ints", C.c_int64), ("bf_vals", C.c_void_p),
                ("n_bf_vals", C.c_int64), ("bf_mant", C.c_double), ("bf_log2", C.c_int32)]


class PlanDesc(C.Structure):
    _fields_ = [("semiring", C.c_int32), ("n_vars", C.c_int32), ("n_checks", C.c_int32), ("n_obs", C.c_int32),
                ("n_steps", C.c_int32), ("w_max", C.c_int32),
                ("hdr", C.POINTER(C.c_int32)), ("ints", C.POINTER(C.c_int32)), ("n_ints", C.c_int64),
                ("tables", C.POINTER(C.c_double)), ("n_tables", C.c_int64),
                ("obs_slot", C.POINTER(C.c_int32)), ("device", C.c_int32), ("sweep", C.POINTER(SweepDesc)),
                ("table_bits", C.c_int32), ("wide", C.POINTER(WideDesc)), ("flags", C.c_int32), ("log2_scale", C.c_int32)]


class ProblemDesc(C.Structure):
    _fields_ = [("semiring", C.c_int32), ("n_vars", C.c_int32), ("n_checks", C.c_int32), ("n_obs", C.c_int32),
                ("n_factors", C.c_int32), ("factor_ptr", C.c_void_p), ("factor_vars", C.c_void_p),
                ("factor_tables", C.c_void_p), ("n_rows", C.c_int32), ("row_ptr", C.c_void_p), ("row_vars", C.c_void_p),
                ("row_kind", C.c_void_p), ("row_index", C.c_void_p), ("order", C.c_void_p), ("head_bits", C.c_int32),
                ("table_bits", C.c_int32), ("device", C.c_int32), ("flags", C.c_int32), ("wide_t_max", C.c_int32)]


COMPILE_NO_SWEEP, COMPILE_NO_FUSE, COMPILE_FORCE_WIDE, COMPILE_DYNAMIC_RESCALE = 1, 2, 4, 8
PLAN_DYNAMIC_RESCALE = 1
(LW_META, LW_COST, LW_ORDER, LW_HDR, LW_INTS, LW_TABLES, LW_OBS_SLOT, LW_SW_REC, LW_SW_TB, LW_SW_LANETAB, LW_SW_TVALS,
 LW_SW_HEAD_BITS, LW_SW_HEAD_STATE, LW_SW_HEAD_CFG, LW_SW_OUT_INDEX, LW_WD_PASS_HDR, LW_WD_STEP_HDR, LW_WD_INTS,
 LW_WD_TABLES, LW_WD_OBS_POS, LW_WD_BF_OFF, LW_WD_BF_INTS, LW_WD_BF_VALS, LW_WD_BF_SCALE) = range(24)
_LW_DTYPE = {LW_COST: np.float64, LW_TABLES: np.float64, LW_SW_TVALS: np.float64, LW_SW_HEAD_STATE: np.float64,
             LW_SW_HEAD_CFG: np.uint64, LW_SW_LANETAB: np.uint32, LW_WD_TABLES: np.float64, LW_WD_BF_VALS: np.float64,
             LW_WD_BF_SCALE: np.float64}


class McDesc(C.Structure):
    _fields_ = [("plan", C.c_void_p), ("H", C.c_void_p), ("L", C.c_void_p), ("row_class", C.POINTER(C.c_int32)),
                ("model", C.c_int32), ("n_sites", C.c_int32),
                ("p0", C.POINTER(C.c_double)), ("p1", C.POINTER(C.c_double)), ("p2", C.POINTER(C.c_double)),
                ("chunk", C.c_int64), ("comm", C.c_void_p)]


_lib = None


def lib():
    """Load the shared library (once).  Raises TqecError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TqecError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        f"or `make -C tensorqec.jl_b200/csrc` (there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    L.tqec_last_error.restype = C.c_char_p
    vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
    L.tqec_device_count.argtypes = [C.POINTER(i32)]
    L.tqec_plan_create.argtypes = [C.POINTER(PlanDesc), C.POINTER(vp)]
    L.tqec_plan_destroy.argtypes = [vp]
    L.tqec_plan_query.argtypes = [vp, i32, C.POINTER(i64)]
    L.tqec_decode_map.argtypes = [vp, vp, i64, vp, vp]
    L.tqec_decode_map_dev.argtypes = [vp, vp, i64, vp, vp, vp]
    L.tqec_decode_marginal.argtypes = [vp, vp, i64, vp, vp]
    L.tqec_decode_marginal_dev.argtypes = [vp, vp, i64, vp, vp, vp]
    L.tqec_gf2_create.argtypes = [i32, i32, vp, i32, C.POINTER(vp)]
    L.tqec_gf2_destroy.argtypes = [vp]
    L.tqec_gf2_apply.argtypes = [vp, vp, i64, vp]
    L.tqec_gf2_apply_dev.argtypes = [vp, vp, i64, vp, vp]
    L.tqec_logical_flags.argtypes = [vp, vp, vp, vp, i64, vp, vp]
    L.tqec_coset_rep.argtypes = [vp, vp, vp, vp, vp, i64, vp, vp]
    L.tqec_sample_errors.argtypes = [i32, i32, vp, vp, vp, u64, i64, i64, vp, i32]
    L.tqec_mc_run.argtypes = [C.POINTER(McDesc), u64, i64, i64, vp, vp]
    L.tqec_fp64_peak.argtypes = [i32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.tqec_lower.argtypes = [C.POINTER(ProblemDesc), C.POINTER(vp)]
    L.tqec_lowered_destroy.argtypes = [vp]
    L.tqec_lowered_get.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(i64)]
    L.tqec_plan_from_lowered.argtypes = [vp, i32, C.POINTER(vp)]
    L.tqec_lowered_save.argtypes = [vp, C.c_char_p]
    L.tqec_lowered_load.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.tqec_plan_compile.argtypes = [C.POINTER(ProblemDesc), C.POINTER(vp)]
    L.tqec_decode_marginal_log2.argtypes = [vp, vp, i64, vp, vp, vp]
    L.tqec_decode_map_bytes.argtypes = [vp, vp, i64, vp, vp]
    L.tqec_decode_map_bytes2.argtypes = [vp, vp, C.c_int32, vp, C.c_int32, i64, vp, vp]
    L.tqec_decode_marginal_bytes.argtypes = [vp, vp, i64, vp, vp]
    L.tqec_table_create.argtypes = [i64, i32, i32, vp, vp, i32, C.POINTER(vp)]
    L.tqec_table_destroy.argtypes = [vp]
    L.tqec_table_decode.argtypes = [vp, vp, i64, vp, vp]
    L.tqec_comm_unique_id.argtypes = [vp]
    L.tqec_comm_init.argtypes = [i32, i32, vp, i32, C.POINTER(vp)]
    L.tqec_comm_destroy.argtypes = [vp]
    L.tqec_comm_allreduce_counts.argtypes = [vp, vp]
    _lib = L
    return L


def check(rc: int):
    if rc != OK:
        msg = lib().tqec_last_error()
        raise TqecError(f"libtqec_cuda error {rc}: {msg.decode() if msg else '?'}")


def device_count() -> int:
    n = C.c_int32(0)
    check(lib().tqec_device_count(C.byref(n)))
    return n.value


def require_device(device: int = 0):
    try:
        n = device_count()
    except TqecError as e:
        raise TqecError(f"no usable CUDA device: {e} (the decoding path has no CPU fallback)") from None
    if n <= device:
        raise TqecError(f"CUDA device {device} requested but {n} visible (the decoding path has no CPU fallback)")


def fp64_peak(device: int = 0):
    """-> dict(dadd_tops, dfma_tflops, maxplus_tops) measured on `device`."""
    require_device(device)
    a, b, c = C.c_double(0), C.c_double(0), C.c_double(0)
    check(lib().tqec_fp64_peak(device, C.byref(a), C.byref(b), C.byref(c)))
    d = C.c_double(0)
    lib().tqec_dmma_peak.argtypes = [C.c_int32, C.POINTER(C.c_double)]
    check(lib().tqec_dmma_peak(device, C.byref(d)))
    return {"dadd_tops": a.value, "dfma_tflops": b.value, "maxplus_tops": c.value, "dmma_tflops": d.value}


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class Problem:
    """A decoding factor graph marshalled for `tqec_lower` / `tqec_plan_compile` (include/tqec.h: tqec_problem_desc).
    factors: objects with .vars / .table (flat, first variable fastest); checks: objects with .vars / .kind ("syn" |
    "obs") / .index -- the same lists the Python lowering takes (schedule.Factor / schedule.Check)."""

    def __init__(self, factors, checks, semiring, n_vars, n_checks, n_obs, order=None, head_bits=0, table_bits=0,
                 device=0, flags=0, wide_t_max=0):
        fptr = np.zeros(len(factors) + 1, dtype=np.int32)
        fv, ft = [], []
        for i, f in enumerate(factors):
            fv += [int(v) for v in f.vars]
            ft.append(np.asarray(f.table, dtype=np.float64).reshape(-1))
            fptr[i + 1] = len(fv)
        rptr = np.zeros(len(checks) + 1, dtype=np.int32)
        rv = []
        for i, c in enumerate(checks):
            rv += [int(v) for v in c.vars]
            rptr[i + 1] = len(rv)
        self._keep = [fptr, _c(fv if fv else [0], np.int32), _c(np.concatenate(ft) if ft else [0.0], np.float64), rptr,
                      _c(rv if rv else [0], np.int32), _c([0 if c.kind == "syn" else 1 for c in checks] or [0], np.int32),
                      _c([c.index for c in checks] or [0], np.int32),
                      _c(order, np.int32) if order is not None else None]
        k = self._keep
        self.desc = ProblemDesc(semiring, n_vars, n_checks, n_obs, len(factors), _ptr(k[0]), _ptr(k[1]), _ptr(k[2]),
                                len(checks), _ptr(k[3]), _ptr(k[4]), _ptr(k[5]), _ptr(k[6]),
                                _ptr(k[7]) if k[7] is not None else None, head_bits, table_bits, device, flags, wide_t_max)


from .mod2 import empty_big as _empty_big  # noqa: E402


class Lowered:
    """Owning handle of a `tqec_lowered`: the C++ lowering's tables (host memory only; works without a GPU)."""

    def __init__(self, problem: Problem):
        h = C.c_void_p()
        check(lib().tqec_lower(C.byref(problem.desc), C.byref(h)))
        self.h = h

    def save(self, path) -> None:
        """Write the lowered plan to `path` (tqec_lowered_save): lower once, create plans from the file afterwards."""
        check(lib().tqec_lowered_save(self.h, os.fsencode(path)))

    @classmethod
    def load(cls, path) -> "Lowered":
        """A lowered plan from a file written by `save` (tqec_lowered_load); refused if the library's table format changed."""
        self = cls.__new__(cls)
        h = C.c_void_p()
        check(lib().tqec_lowered_load(os.fsencode(path), C.byref(h)))
        self.h = h
        return self

    def get(self, what: int) -> np.ndarray:
        ptr, n = C.c_void_p(), C.c_int64(0)
        check(lib().tqec_lowered_get(self.h, what, C.byref(ptr), C.byref(n)))
        dt = np.dtype(_LW_DTYPE.get(what, np.int32))
        if n.value == 0:
            return np.zeros(0, dtype=dt)
        buf = (C.c_char * (n.value * dt.itemsize)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=dt).copy()

    @property
    def meta(self):
        m = self.get(LW_META)
        return dict(kind=int(m[0]), n_steps=int(m[1]), w_max=int(m[2]), log2_scale=int(m[3]), W=int(m[4]), sg=int(m[5]),
                    n_ss=int(m[6]), n_head_bits=int(m[7]), bp_words=int(m[8]), head_steps=int(m[9]), conflicts=int(m[10]),
                    n_pass=int(m[11]), wide_steps=int(m[12]), w_cap=int(m[13]), t_max=int(m[14]), table_bits=int(m[15]))

    def close(self):
        if getattr(self, "h", None):
            lib().tqec_lowered_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Plan:
    """Owning handle of a `tqec_plan`."""

    @classmethod
    def compile(cls, problem: "Problem"):
        """Factor graph -> plan through the library's own lowering (tqec_lower + tqec_plan_from_lowered): the path a
        non-Python host takes with tqec_plan_compile.  The lowering's summary stays available as `plan.lowered`."""
        d = problem.desc
        require_device(d.device)
        self = cls.__new__(cls)
        self.sch = None
        self.device = d.device
        self.n_obs = d.n_obs
        self.nsw = max(1, (d.n_checks + 63) // 64)
        self.ncw = max(1, (d.n_vars + 63) // 64)
        lw = Lowered(problem)
        self.lowered = lw.meta
        self.order = [int(i) for i in lw.get(LW_ORDER)]
        self.cost = [float(x) for x in lw.get(LW_COST)]
        h = C.c_void_p()
        check(lib().tqec_plan_from_lowered(lw.h, d.device, C.byref(h)))
        lw.close()
        self.h = h
        return self

    @classmethod
    def from_lowered(cls, lw: "Lowered", device: int = 0):
        """Plan from a lowered plan (e.g. `Lowered.load(path)`): no lowering, only the upload (tqec_plan_from_lowered)."""
        require_device(device)
        self = cls.__new__(cls)
        self.sch = None
        self.device = device
        m = lw.meta
        dims = lw.get(LW_META)
        self.lowered = m
        self.n_obs, n_checks, n_vars = int(dims[16]), int(dims[17]), int(dims[18])
        self.nsw = max(1, (n_checks + 63) // 64)
        self.ncw = max(1, (n_vars + 63) // 64)
        self.order = [int(i) for i in lw.get(LW_ORDER)]
        self.cost = [float(x) for x in lw.get(LW_COST)]
        h = C.c_void_p()
        check(lib().tqec_plan_from_lowered(lw.h, device, C.byref(h)))
        self.h = h
        return self

    def __init__(self, sch, device: int = 0):
        require_device(device)
        self.sch = sch
        self.device = device
        self.n_obs = sch.n_obs
        self.nsw = max(1, (sch.n_checks + 63) // 64)
        self.ncw = max(1, (sch.n_vars + 63) // 64)
        tb = getattr(sch, "table_bits", None)
        table_bits = 0 if tb is None else (-1 if int(tb) <= 0 else int(tb))   # ABI: 0 = library default, -1 = never
        if hasattr(sch, "pass_hdr"):
            # global-memory lowering (wide.py): the descriptor carries only the `wide` tables
            keep = [_c(sch.pass_hdr, np.int32), _c(sch.step_hdr, np.int32), _c(sch.ints, np.int32),
                    _c(sch.tables, np.float64), _c(sch.obs_pos if sch.obs_pos else [0], np.int32)]
            wd = WideDesc(len(sch.passes), keep[1].shape[0], sch.w_cap, sch.t_max, _ptr(keep[0]), _ptr(keep[1]), _ptr(keep[2]),
                          keep[2].size, _ptr(keep[3]), keep[3].size, _ptr(keep[4]))
            d = PlanDesc(sch.semiring, sch.n_vars, sch.n_checks, sch.n_obs, 0, sch.w_cap, None, None, 0, None, 0, None,
                         device, None, table_bits, C.pointer(wd), int(getattr(sch, "plan_flags", 0)), int(sch.log2_scale))
            h = C.c_void_p()
            check(lib().tqec_plan_create(C.byref(d), C.byref(h)))
            self.h = h
            return
        hdr = _c(sch.hdr, np.int32)
        ints = _c(sch.ints, np.int32)
        tabs = _c(sch.tables, np.float64)
        obs = _c(sch.obs_slot if sch.obs_slot else [0], np.int32)
        d = PlanDesc(sch.semiring, sch.n_vars, sch.n_checks, sch.n_obs, len(sch.steps), sch.w_max,
                     hdr.ctypes.data_as(C.POINTER(C.c_int32)), ints.ctypes.data_as(C.POINTER(C.c_int32)), ints.size,
                     tabs.ctypes.data_as(C.POINTER(C.c_double)), tabs.size,
                     obs.ctypes.data_as(C.POINTER(C.c_int32)), device, None, table_bits, None, 0,
                     int(getattr(sch, "log2_scale", 0) or 0))
        sw = getattr(sch, "sweep", None)
        if sw is not None:
            keep = [_c(sw.rec, np.int32), _c(sw.tb, np.int32), _c(sw.lanetab, np.uint32), _c(sw.tvals, np.float64),
                    _c(sw.head_bits if sw.head_bits else [0], np.int32), _c(sw.head_state, np.float64),
                    _c(sw.head_cfg, np.uint64), _c(sw.out_index, np.int32)]
            sd = SweepDesc(sw.W, sw.sg, len(sw.ssteps), len(sw.head_bits), sw.bp_words, keep[3].size,
                           *[_ptr(a) for a in keep])
            d.sweep = C.pointer(sd)
        h = C.c_void_p()
        check(lib().tqec_plan_create(C.byref(d), C.byref(h)))
        self.h = h

    def query(self, what: int) -> int:
        v = C.c_int64(0)
        check(lib().tqec_plan_query(self.h, what, C.byref(v)))
        return v.value

    def geometry(self):
        return {k: self.query(q) for k, q in [("team_threads", Q_TEAM_THREADS), ("shots_per_team", Q_SHOTS_PER_TEAM),
                                               ("smem_bytes", Q_SMEM_BYTES), ("grid", Q_GRID),
                                               ("teams_per_sm", Q_TEAMS_PER_SM), ("bp_bytes_per_team", Q_BP_BYTES_PER_TEAM),
                                               ("candidates_per_shot", Q_CANDIDATES_PER_SHOT), ("sm_count", Q_SM_COUNT),
                                               ("sweep", Q_SWEEP), ("table", Q_TABLE), ("wide", Q_WIDE)]}

    def decode_map(self, synd_words: np.ndarray, want_logp=True):
        s = _c(synd_words, np.uint64).reshape(-1, self.nsw)
        B = s.shape[0]
        corr = np.zeros((B, self.ncw), dtype=np.uint64)
        logp = np.zeros(B, dtype=np.float64) if want_logp else None
        check(lib().tqec_decode_map(self.h, _ptr(s), B, _ptr(corr), _ptr(logp) if want_logp else None))
        return corr, logp

    def decode_map_bits(self, synd_bits: np.ndarray, n_vars: int):
        """(B, n_checks) uint8 0/1 -> ((B, n_vars) uint8, logp): one byte per bit both ways, packed on the device."""
        s = _c(synd_bits, np.uint8)
        B = s.shape[0]
        corr = _empty_big((B, n_vars), np.uint8)
        logp = _empty_big(B, np.float64)
        check(lib().tqec_decode_map_bytes(self.h, _ptr(s), B, _ptr(corr), _ptr(logp)))
        return corr, logp

    def decode_map_bits2(self, sa: np.ndarray, sb: np.ndarray, n_vars: int):
        """The same with the syndrome as two (B, n) uint8 arrays (a CSS code's sx and sz): no concatenation on the host."""
        a, b = _c(sa, np.uint8), _c(sb, np.uint8)
        B = a.shape[0]
        corr = _empty_big((B, n_vars), np.uint8)
        logp = _empty_big(B, np.float64)
        check(lib().tqec_decode_map_bytes2(self.h, _ptr(a), a.shape[1], _ptr(b), b.shape[1], B, _ptr(corr), _ptr(logp)))
        return corr, logp

    def decode_marginal_bits(self, synd_bits: np.ndarray):
        s = _c(synd_bits, np.uint8)
        B = s.shape[0]
        mar = _empty_big((B, 1 << self.n_obs), np.float64)
        arg = _empty_big(B, np.int32)
        check(lib().tqec_decode_marginal_bytes(self.h, _ptr(s), B, _ptr(mar), _ptr(arg)))
        return mar, arg

    def decode_map_dev(self, d_synd: int, B: int, d_corr: int, d_logp: int = 0, stream: int = 0):
        check(lib().tqec_decode_map_dev(self.h, d_synd, B, d_corr, d_logp or None, stream or None))

    def decode_marginal(self, synd_words: np.ndarray):
        s = _c(synd_words, np.uint64).reshape(-1, self.nsw)
        B = s.shape[0]
        mar = np.zeros((B, 1 << self.n_obs), dtype=np.float64)
        arg = np.zeros(B, dtype=np.int32)
        check(lib().tqec_decode_marginal(self.h, _ptr(s), B, _ptr(mar), _ptr(arg)))
        return mar, arg                                    # (the library has undone the static scaling of the tables)

    def decode_marginal_log2(self, synd_words: np.ndarray):
        """-> (mantissas (B, 2^n_obs), log2 (B,), argmax): true marginal = mantissa * 2^log2 (tqec_decode_marginal_log2)."""
        s = _c(synd_words, np.uint64).reshape(-1, self.nsw)
        B = s.shape[0]
        mar = np.zeros((B, 1 << self.n_obs), dtype=np.float64)
        lg = np.zeros(B, dtype=np.int32)
        arg = np.zeros(B, dtype=np.int32)
        check(lib().tqec_decode_marginal_log2(self.h, _ptr(s), B, _ptr(mar), _ptr(lg), _ptr(arg)))
        return mar, lg, arg

    def decode_marginal_dev(self, d_synd: int, B: int, d_mar: int, d_argmax: int = 0, stream: int = 0):
        check(lib().tqec_decode_marginal_dev(self.h, d_synd, B, d_mar, d_argmax or None, stream or None))

    def close(self):
        if getattr(self, "h", None):
            lib().tqec_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GF2Matrix:
    """Owning handle of a `tqec_gf2` (bit-packed rows on the device)."""

    def __init__(self, M: np.ndarray, device: int = 0):
        from .mod2 import pack_rows
        require_device(device)
        M = np.asarray(M, dtype=np.uint8)
        if M.ndim != 2:
            raise ValueError("GF2Matrix expects a 2-D 0/1 matrix")
        self.rows, self.cols = M.shape
        self.rw = max(1, (self.rows + 63) // 64)
        self.cw = max(1, (self.cols + 63) // 64)
        packed = pack_rows(M) if self.rows else np.zeros((0, self.cw), dtype=np.uint64)
        h = C.c_void_p()
        check(lib().tqec_gf2_create(self.rows, self.cols, _ptr(packed) if self.rows else None, device, C.byref(h)))
        self.h = h
        self.device = device

    def apply(self, words: np.ndarray) -> np.ndarray:
        x = _c(words, np.uint64).reshape(-1, self.cw)
        out = np.zeros((x.shape[0], self.rw), dtype=np.uint64)
        check(lib().tqec_gf2_apply(self.h, _ptr(x), x.shape[0], _ptr(out)))
        return out

    def apply_dev(self, d_in: int, B: int, d_out: int, stream: int = 0):
        check(lib().tqec_gf2_apply_dev(self.h, d_in, B, d_out, stream or None))

    def logical_flags(self, row_class, e1_words, e2_words=None):
        e1 = _c(e1_words, np.uint64).reshape(-1, self.cw)
        B = e1.shape[0]
        e2 = None if e2_words is None else _c(e2_words, np.uint64).reshape(-1, self.cw)
        if e2 is not None and e2.shape != e1.shape:
            raise ValueError("error patterns must have the same shape")
        cls = _c(row_class, np.int32)
        if cls.size != self.rows:
            raise ValueError("row_class needs one entry per logical row")
        flags = np.zeros(B, dtype=np.uint8)
        counts = np.zeros(4, dtype=np.int64)
        check(lib().tqec_logical_flags(self.h, _ptr(cls), _ptr(e1), _ptr(e2) if e2 is not None else None, B,
                                       _ptr(flags), _ptr(counts)))
        return flags, counts

    def close(self):
        if getattr(self, "h", None):
            lib().tqec_gf2_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def coset_rep(R: GF2Matrix, L, FIX, synd_words, sector):
    """-> (error words, ok): ok[b] = the pattern lies in the requested sector."""
    s = _c(synd_words, np.uint64).reshape(-1, R.cw)
    B = s.shape[0]
    out = np.zeros((B, R.rw), dtype=np.uint64)
    ok = np.ones(B, dtype=np.uint8)
    sec = _c(sector, np.int32)
    check(lib().tqec_coset_rep(R.h, L.h if L is not None else None, FIX.h if FIX is not None else None, _ptr(s),
                               _ptr(sec), B, _ptr(out), _ptr(ok)))
    return out, ok.astype(bool)


def sample_errors(model: int, probs, seed: int, shot_offset: int, B: int, device: int = 0) -> np.ndarray:
    require_device(device)
    ps = [_c(p, np.float64) for p in probs]
    n = ps[0].size
    nbits = 2 * n if model == MODEL_DEPOL else n
    out = np.zeros((B, max(1, (nbits + 63) // 64)), dtype=np.uint64)
    p1 = _ptr(ps[1]) if model == MODEL_DEPOL else None
    p2 = _ptr(ps[2]) if model == MODEL_DEPOL else None
    check(lib().tqec_sample_errors(model, n, _ptr(ps[0]), p1, p2, C.c_uint64(seed & (2 ** 64 - 1)), shot_offset, B,
                                   _ptr(out), device))
    return out


class Table:
    """Owning handle of a `tqec_table` (sorted syndrome keys -> error words, resident on one device)."""

    def __init__(self, keys: np.ndarray, values: np.ndarray, n_checks: int, n_vars: int, device: int = 0):
        require_device(device)
        self.nsw, self.ncw = max(1, (n_checks + 63) // 64), max(1, (n_vars + 63) // 64)
        k = _c(keys, np.uint64).reshape(-1, self.nsw)
        v = _c(values, np.uint64).reshape(-1, self.ncw)
        if k.shape[0] != v.shape[0]:
            raise ValueError("one value per key")
        h = C.c_void_p()
        check(lib().tqec_table_create(k.shape[0], n_checks, n_vars, _ptr(k), _ptr(v), device, C.byref(h)))
        self.h = h

    def decode(self, synd_words: np.ndarray):
        s = _c(synd_words, np.uint64).reshape(-1, self.nsw)
        B = s.shape[0]
        corr = np.zeros((B, self.ncw), dtype=np.uint64)
        found = np.zeros(B, dtype=np.uint8)
        check(lib().tqec_table_decode(self.h, _ptr(s), B, _ptr(corr), _ptr(found)))
        return corr, found.astype(bool)

    def close(self):
        if getattr(self, "h", None):
            lib().tqec_table_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Comm:
    """Owning handle of a `tqec_comm`: this rank's NCCL communicator inside the library.  `unique_id()` on rank 0, ship
    the 128 bytes to the other ranks, then `Comm(nranks, rank, id, device)` on every rank."""

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_char * 128)()
        check(lib().tqec_comm_unique_id(buf))
        return bytes(buf)

    def __init__(self, nranks: int, rank: int, unique_id: bytes, device: int = 0):
        require_device(device)
        if len(unique_id) != 128:
            raise ValueError("an NCCL unique id has 128 bytes")
        h = C.c_void_p()
        check(lib().tqec_comm_init(nranks, rank, C.c_char_p(unique_id), device, C.byref(h)))
        self.h, self.nranks, self.rank, self.device = h, nranks, rank, device

    def allreduce_counts(self, counts) -> np.ndarray:
        c = _c(counts, np.int64).copy()
        if c.size != 4:
            raise ValueError("four counters")
        check(lib().tqec_comm_allreduce_counts(self.h, _ptr(c)))
        return c

    def close(self):
        if getattr(self, "h", None):
            lib().tqec_comm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def mc_run(plan: Plan, H: GF2Matrix, L: GF2Matrix, row_class, model: int, probs, seed: int, shot_offset: int,
           n_shots: int, chunk: int = 0, comm: "Comm" = None):
    ps = [_c(p, np.float64) for p in probs]
    cls = _c(row_class, np.int32)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    d = McDesc(plan.h, H.h, L.h, cls.ctypes.data_as(C.POINTER(C.c_int32)), model, ps[0].size, dp(ps[0]),
               dp(ps[1]) if model == MODEL_DEPOL else None, dp(ps[2]) if model == MODEL_DEPOL else None, chunk,
               comm.h if comm is not None else None)
    counts = np.zeros(4, dtype=np.int64)
    ms = C.c_float(0.0)
    check(lib().tqec_mc_run(C.byref(d), C.c_uint64(seed & (2 ** 64 - 1)), shot_offset, n_shots, _ptr(counts),
                            C.byref(ms)))
    return counts, ms.value
