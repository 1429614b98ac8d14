"""Belief propagation + OSD decoder, batched on the GPU: a baseline sharing the GF(2) front and back end of the tensor-
network decoders.

Reference: src/decoding/bposd.jl -- `BPDecoder(bp_max_iter = 100, osd = true)` (:16-19), `compile` (:30-35:
mu_i = log((1 - p_i) / p_i)), `belief_propagation` (:55-78), `osd` (:80-97), `decode` (:45-52: BP's pattern when it
reproduces the syndrome, else the OSD pattern; without OSD a failed BP returns the zero pattern with success_tag false).
The kernel (`k_bp_osd`, csrc/tqec_bp.cu) runs one shot per thread with the messages in [edge][shot] layout.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _cabi
from .error_model import IndependentFlipError, SimpleSyndrome
from .mod2 import as_bits, pack_bits, unpack_bits
from .tanner import SimpleTannerGraph


@dataclass
class BPDecoder:
    """bposd.jl:16-19."""
    bp_max_iter: int = 100
    osd: bool = True
    device: int = 0


class CompiledBP:
    def __init__(self, tanner: SimpleTannerGraph, p, bp_max_iter: int, osd: bool, device: int = 0):
        _cabi.require_device(device)
        self.tanner, self.bp_max_iter, self.osd = tanner, bp_max_iter, osd
        s_ptr = np.zeros(tanner.ns + 1, dtype=np.int32)
        adj = []
        for s, qs in enumerate(tanner.s2q):
            adj += [int(q) for q in qs]
            s_ptr[s + 1] = len(adj)
        adj = np.asarray(adj, dtype=np.int32)
        p = np.ascontiguousarray(p, dtype=np.float64)
        lib = _cabi.lib()
        lib.tqec_bp_create.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                       C.POINTER(C.c_void_p)]
        lib.tqec_bp_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        lib.tqec_bp_destroy.argtypes = [C.c_void_p]
        h = C.c_void_p()
        _cabi.check(lib.tqec_bp_create(tanner.nq, tanner.ns, s_ptr.ctypes.data_as(C.c_void_p), adj.ctypes.data_as(C.c_void_p),
                                       p.ctypes.data_as(C.c_void_p), bp_max_iter, int(osd), device, C.byref(h)))
        self.h = h

    def decode_words(self, synd_words: np.ndarray):
        s = np.ascontiguousarray(synd_words, dtype=np.uint64).reshape(-1, max(1, (self.tanner.ns + 63) // 64))
        B = s.shape[0]
        corr = np.zeros((B, max(1, (self.tanner.nq + 63) // 64)), dtype=np.uint64)
        flags = np.zeros(B, dtype=np.uint8)
        _cabi.check(_cabi.lib().tqec_bp_decode(self.h, s.ctypes.data_as(C.c_void_p), B, corr.ctypes.data_as(C.c_void_p),
                                               flags.ctypes.data_as(C.c_void_p)))
        return corr, flags

    def __del__(self):
        try:
            if getattr(self, "h", None):
                _cabi.lib().tqec_bp_destroy(self.h)
                self.h = None
        except Exception:
            pass


def compile_bp(decoder: BPDecoder, problem) -> CompiledBP:
    """bposd.jl:30-35 (problem: ClassicalDecodingProblem)."""
    return CompiledBP(problem.tanner, problem.pvec.p, decoder.bp_max_iter, decoder.osd, decoder.device)


def decode_bp(cb: CompiledBP, syndrome: SimpleSyndrome):
    """bposd.jl:45-52, batched."""
    from .decoding import DecodingResult
    bits = as_bits(syndrome.s)
    single = bits.ndim == 1
    bits = np.atleast_2d(bits)
    if bits.shape[1] != cb.tanner.ns:
        raise ValueError(f"syndrome has {bits.shape[1]} bits, the decoder expects {cb.tanner.ns}")
    corr, flags = cb.decode_words(pack_bits(bits))
    e = unpack_bits(corr, cb.tanner.nq)
    ok = flags != 0
    if single:
        return DecodingResult(bool(ok[0]), e[0])
    return DecodingResult(ok, e)
