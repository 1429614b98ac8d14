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
                ("tables", C.c_void_p), ("n_tables", C.c_int64), ("obs_pos", C.c_void_p)]


class PlanDesc(C.Structure):
    _fields_ = [("semiring", C.c_int32), ("n_vars", C.c_int32), ("n_checks", C.c_int32), ("n_obs", C.c_int32),
                ("n_steps", C.c_int32), ("w_max", C.c_int32),
                ("hdr", C.POINTER(C.c_int32)), ("ints", C.POINTER(C.c_int32)), ("n_ints", C.c_int64),
                ("tables", C.POINTER(C.c_double)), ("n_tables", C.c_int64),
                ("obs_slot", C.POINTER(C.c_int32)), ("device", C.c_int32), ("sweep", C.POINTER(SweepDesc)),
                ("table_bits", C.c_int32), ("wide", C.POINTER(WideDesc))]


class McDesc(C.Structure):
    _fields_ = [("plan", C.c_void_p), ("H", C.c_void_p), ("L", C.c_void_p), ("row_class", C.POINTER(C.c_int32)),
                ("model", C.c_int32), ("n_sites", C.c_int32),
                ("p0", C.POINTER(C.c_double)), ("p1", C.POINTER(C.c_double)), ("p2", C.POINTER(C.c_double)),
                ("chunk", C.c_int64)]


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
    return {"dadd_tops": a.value, "dfma_tflops": b.value, "maxplus_tops": c.value}


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class Plan:
    """Owning handle of a `tqec_plan`."""

    def __init__(self, sch, device: int = 0):
        require_device(device)
        self.sch = sch
        self.device = device
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
                         device, None, table_bits, C.pointer(wd))
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
                     obs.ctypes.data_as(C.POINTER(C.c_int32)), device, None, table_bits, None)
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

    def decode_map_dev(self, d_synd: int, B: int, d_corr: int, d_logp: int = 0, stream: int = 0):
        check(lib().tqec_decode_map_dev(self.h, d_synd, B, d_corr, d_logp or None, stream or None))

    def decode_marginal(self, synd_words: np.ndarray):
        s = _c(synd_words, np.uint64).reshape(-1, self.nsw)
        B = s.shape[0]
        mar = np.zeros((B, 1 << self.sch.n_obs), dtype=np.float64)
        arg = np.zeros(B, dtype=np.int32)
        check(lib().tqec_decode_marginal(self.h, _ptr(s), B, _ptr(mar), _ptr(arg)))
        if self.sch.log2_scale:
            mar = np.ldexp(mar, self.sch.log2_scale)       # undo the static power-of-two scaling of the factor tables
        return mar, arg

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


def mc_run(plan: Plan, H: GF2Matrix, L: GF2Matrix, row_class, model: int, probs, seed: int, shot_offset: int,
           n_shots: int, chunk: int = 0):
    ps = [_c(p, np.float64) for p in probs]
    cls = _c(row_class, np.int32)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    d = McDesc(plan.h, H.h, L.h, cls.ctypes.data_as(C.POINTER(C.c_int32)), model, ps[0].size, dp(ps[0]),
               dp(ps[1]) if model == MODEL_DEPOL else None, dp(ps[2]) if model == MODEL_DEPOL else None, chunk)
    counts = np.zeros(4, dtype=np.int64)
    ms = C.c_float(0.0)
    check(lib().tqec_mc_run(C.byref(d), C.c_uint64(seed & (2 ** 64 - 1)), shot_offset, n_shots, _ptr(counts),
                            C.byref(ms)))
    return counts, ms.value
