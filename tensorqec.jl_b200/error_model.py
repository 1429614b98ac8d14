"""Error models, error patterns, syndromes and the GF(2) operations around the decoder.

Reference: src/decoding/error_model.jl:13-15 (`IndependentFlipError`), :41-45 (`IndependentDepolarizingError`),
:28-29 / :61-62 (`iid_error`), :69-71 / :97-117 (`random_error_pattern`), :81-84 (`CSSErrorPattern`),
:119-122 (`SimpleSyndrome`), :138-142 (`CSSSyndrome`), :131-136 / :144-146 (`syndrome_extraction`),
:161-163 / :179-181 (`check_logical_error`).

Every operation runs on the GPU through the C ABI (bit-packed popcount-parity kernels); containers accept one shot
(1-D arrays) or a batch (2-D arrays, shots first).  Sampling uses a counter-based Philox stream instead of Julia's
task-local RNG (not reproducible outside Julia): `random_error_pattern(em; seed, shots, shot_offset)`.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _cabi
from .mod2 import as_bits, pack_bits, unpack_bits, validated
from .tanner import CSSTannerGraph, SimpleTannerGraph


class AbstractErrorModel:
    pass


class AbstractClassicalErrorModel(AbstractErrorModel):
    pass


class AbstractQuantumErrorModel(AbstractErrorModel):
    pass


@dataclass
class IndependentFlipError(AbstractClassicalErrorModel):
    p: np.ndarray

    def __post_init__(self):
        self.p = np.asarray(self.p, dtype=np.float64)


@dataclass
class IndependentDepolarizingError(AbstractQuantumErrorModel):
    px: np.ndarray
    py: np.ndarray
    pz: np.ndarray

    def __post_init__(self):
        self.px, self.py, self.pz = (np.asarray(a, dtype=np.float64) for a in (self.px, self.py, self.pz))
        if not (self.px.shape == self.py.shape == self.pz.shape):
            raise ValueError("px, py, pz must have the same length")


def iid_error(*args):
    """iid_error(p, n) | iid_error(p, tanner::SimpleTannerGraph) -> IndependentFlipError;
    iid_error(px, py, pz, n) | iid_error(p, tanner::CSSTannerGraph) -> IndependentDepolarizingError
    (error_model.jl:28-29, 61-62: a CSS graph gets (p, p, p))."""
    if len(args) == 2:
        p, n = args
        if isinstance(n, CSSTannerGraph):
            k = n.stgx.nq
            return IndependentDepolarizingError(np.full(k, p), np.full(k, p), np.full(k, p))
        if isinstance(n, SimpleTannerGraph):
            n = n.nq
        return IndependentFlipError(np.full(int(n), p))
    if len(args) == 4:
        px, py, pz, n = args
        return IndependentDepolarizingError(np.full(int(n), px), np.full(int(n), py), np.full(int(n), pz))
    raise TypeError("iid_error(p, n|tanner) or iid_error(px, py, pz, n)")


@dataclass
class CSSErrorPattern:
    xerror: np.ndarray
    zerror: np.ndarray

    def __post_init__(self):
        self.xerror, self.zerror = as_bits(self.xerror), as_bits(self.zerror)


@dataclass
class SimpleSyndrome:
    s: np.ndarray

    def __post_init__(self):
        self.s = validated(self.s)                            # checked once here; decode() does not scan the batch again

    def __eq__(self, o):
        return isinstance(o, SimpleSyndrome) and np.array_equal(self.s, o.s)


@dataclass
class CSSSyndrome:
    sx: np.ndarray
    sz: np.ndarray

    def __post_init__(self):
        self.sx, self.sz = validated(self.sx), validated(self.sz)

    def __eq__(self, o):
        return isinstance(o, CSSSyndrome) and np.array_equal(self.sx, o.sx) and np.array_equal(self.sz, o.sz)


# ------------------------------------------------------------------------------------------------------------
def random_error_pattern(em: AbstractErrorModel, *, seed: int = 0, shots=None, shot_offset: int = 0, device: int = 0):
    """error_model.jl:69-71 (flip: u < p[i]) and :97-117 (depolarizing: one uniform per qubit, Y tested first).
    `shots=None` -> one pattern (1-D arrays); otherwise a batch of `shots` patterns for shot indices
    shot_offset .. shot_offset+shots-1 of the Philox stream `seed`."""
    from .dem import DetectorErrorModel
    if isinstance(em, DetectorErrorModel):                  # dem.jl:162-164
        em = IndependentFlipError(em.error_rates)
    B = 1 if shots is None else int(shots)
    if isinstance(em, IndependentFlipError):
        w = _cabi.sample_errors(_cabi.MODEL_FLIP, [em.p], seed, shot_offset, B, device)
        bits = unpack_bits(w, em.p.size)
        return bits[0] if shots is None else bits
    if isinstance(em, IndependentDepolarizingError):
        n = em.px.size
        w = _cabi.sample_errors(_cabi.MODEL_DEPOL, [em.px, em.py, em.pz], seed, shot_offset, B, device)
        bits = unpack_bits(w, 2 * n)
        x, z = bits[:, :n], bits[:, n:]
        return CSSErrorPattern(x[0], z[0]) if shots is None else CSSErrorPattern(x, z)
    raise TypeError(f"unsupported error model {type(em).__name__}")


_MAT_CACHE = {}


def _device_matrix(M: np.ndarray, device: int) -> "_cabi.GF2Matrix":
    M = np.ascontiguousarray(M, dtype=np.uint8)
    key = (M.shape, M.tobytes(), device)
    m = _MAT_CACHE.get(key)
    if m is None:
        if len(_MAT_CACHE) > 64:
            _MAT_CACHE.clear()
        m = _MAT_CACHE[key] = _cabi.GF2Matrix(M, device)
    return m


def _apply(H: np.ndarray, e: np.ndarray, device: int) -> np.ndarray:
    e = as_bits(e)
    single = e.ndim == 1
    e2 = e[None, :] if single else e
    if e2.shape[1] != H.shape[1]:
        raise ValueError(f"error pattern has {e2.shape[1]} bits, the check matrix has {H.shape[1]} columns")  # DimensionMismatch
    out = unpack_bits(_device_matrix(H, device).apply(pack_bits(e2)), H.shape[0])
    return out[0] if single else out


def syndrome_extraction(errored, H_or_tanner, *, device: int = 0):
    """error_model.jl:131-136: s = H e;  :144-146: CSS -> CSSSyndrome(sx = Hx ez, sz = Hz ex)."""
    if isinstance(errored, CSSErrorPattern):
        t = H_or_tanner
        if not isinstance(t, CSSTannerGraph):
            raise TypeError("a CSSErrorPattern needs a CSSTannerGraph")
        return CSSSyndrome(_apply(t.stgx.H, errored.zerror, device), _apply(t.stgz.H, errored.xerror, device))
    H = H_or_tanner.H if isinstance(H_or_tanner, SimpleTannerGraph) else as_bits(H_or_tanner)
    return SimpleSyndrome(_apply(H, errored, device))


def check_logical_error(e1, e2, *logicals, device: int = 0):
    """check_logical_error(e1, e2, lz) (error_model.jl:161-163) -> any_i lz[i].(e1 - e2);
    check_logical_error(ep1::CSSErrorPattern, ep2, lx, lz) (:179-181) -> check(z1, z2, lx) || check(x1, x2, lz)."""
    if isinstance(e1, CSSErrorPattern):
        lx, lz = (as_bits(l) for l in logicals)
        n = e1.xerror.shape[-1]
        single = e1.xerror.ndim == 1
        a = np.concatenate([np.atleast_2d(e1.xerror), np.atleast_2d(e1.zerror)], axis=1)
        b = np.concatenate([np.atleast_2d(e2.xerror), np.atleast_2d(e2.zerror)], axis=1)
        L = np.zeros((lx.shape[0] + lz.shape[0], 2 * n), dtype=np.uint8)
        L[: lz.shape[0], :n] = lz                       # class 0: lz rows on the x block
        L[lz.shape[0]:, n:] = lx                        # class 1: lx rows on the z block
        cls = [0] * lz.shape[0] + [1] * lx.shape[0]
        flags, _ = _device_matrix(L, device).logical_flags(cls, pack_bits(a), pack_bits(b))
        res = flags != 0
        return bool(res[0]) if single else res
    (L,) = logicals
    L = as_bits(L)
    a, b = as_bits(e1), as_bits(e2)
    single = a.ndim == 1
    flags, _ = _device_matrix(L, device).logical_flags([0] * L.shape[0], pack_bits(np.atleast_2d(a)),
                                                      pack_bits(np.atleast_2d(b)))
    res = flags != 0
    return bool(res[0]) if single else res
