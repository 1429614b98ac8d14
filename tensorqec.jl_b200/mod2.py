"""GF(2) host helpers: the `Mod2` scalar of the reference and the bit-packed layouts the C-ABI uses.

Reference: src/codes/mod2.jl:20-41 (Mod2 algebra: + and - are XOR, * is AND), :44-71 (`bitmul!`,
`compresscol`: 64 rows per UInt64, bit k of word w = row 64*w + k).

Host containers are numpy uint8 vectors / matrices holding 0/1 (the reference's `Vector{Mod2}` is one byte
per bit as well).  Everything that crosses the C-ABI is bit-packed, SHOT-MAJOR:

    packed[shot, w] (uint64), bit k of word w  <->  bit index 64*w + k          (little-endian bits)

which is `compresscol` applied to the (bits x shots) matrix.
"""
from __future__ import annotations

import numpy as np


class Mod2:
    """GF(2) scalar (src/codes/mod2.jl:20-41).  Kept for API parity; bulk data uses uint8 arrays."""

    __slots__ = ("x",)

    def __init__(self, x):
        if isinstance(x, Mod2):
            x = x.x
        if x not in (0, 1, True, False):
            raise ValueError(f"Mod2 expects a Bool or 0/1, got {x!r}")  # Bool(x) InexactError in the reference
        self.x = bool(x)

    def __add__(self, o):
        return Mod2(self.x ^ Mod2(o).x)

    __sub__ = __add__
    __radd__ = __add__

    def __neg__(self):
        return self

    def __mul__(self, o):
        return Mod2(self.x and Mod2(o).x)

    def __eq__(self, o):
        return isinstance(o, (Mod2, bool, int)) and self.x == Mod2(o).x

    def __hash__(self):
        return hash(self.x)

    def __int__(self):
        return 1 if self.x else 0

    def __bool__(self):
        return self.x

    def iszero(self):
        return not self.x

    def __repr__(self):
        return "1₂" if self.x else "0₂"


def as_bits(v) -> np.ndarray:
    """Anything bit-like (list of 0/1, bools, Mod2) -> uint8 array of 0/1."""
    if getattr(v, "_tqec_validated", False):                    # already checked (ValidatedBits)
        return v
    if isinstance(v, np.ndarray) and v.dtype != object:
        a = v.astype(np.uint8, copy=False)
    else:
        a = np.array([[int(x) for x in row] if isinstance(row, (list, tuple, np.ndarray)) else int(row) for row in v],
                     dtype=np.uint8) if len(v) else np.zeros(0, dtype=np.uint8)
    if a.size and _has_non_bit(a):
        raise ValueError("bits must be 0/1")
    return a


class ValidatedBits(np.ndarray):
    """uint8 0/1 array that already went through as_bits, or that the library produced itself (decoded corrections):
    as_bits returns it as it is instead of scanning a large batch again.  Views and slices keep the mark."""
    _tqec_validated = True

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        # arithmetic on validated bits gives plain arrays (the result need not be 0/1); only views and slices keep the mark
        plain = tuple(np.asarray(x) if isinstance(x, ValidatedBits) else x for x in inputs)
        if out is not None:
            kwargs["out"] = tuple(np.asarray(x) if isinstance(x, ValidatedBits) else x for x in out)
        return getattr(ufunc, method)(*plain, **kwargs)


def validated(v) -> np.ndarray:
    """as_bits + the ValidatedBits mark: containers that checked their bits once (SimpleSyndrome, CSSSyndrome) hand them
    to `decode` without a second pass over a batch-sized array."""
    a = as_bits(v)
    return a if isinstance(a, ValidatedBits) else a.view(ValidatedBits)


_LIBC = None


def empty_big(shape, dtype) -> np.ndarray:
    """np.empty for batch-sized arrays; large ones ask for transparent huge pages (MADV_HUGEPAGE) before their first
    touch -- faulting in 324 MB of corrections (2e6 shots at d = 9) in 4 KiB pages costs more than decoding them."""
    global _LIBC
    a = np.empty(shape, dtype=dtype)
    if a.nbytes >= (32 << 20):
        try:
            import ctypes
            if _LIBC is None:
                _LIBC = ctypes.CDLL("libc.so.6", use_errno=True)
            addr = a.ctypes.data
            lo = (addr + (1 << 21) - 1) & ~((1 << 21) - 1)
            ln = (addr + a.nbytes - lo) & ~((1 << 21) - 1)
            if ln > 0:
                _LIBC.madvise(ctypes.c_void_p(lo), ctypes.c_size_t(ln), 14)   # MADV_HUGEPAGE; a refusal is harmless
        except Exception:                                        # noqa: BLE001
            pass
    return a


def concat_bits(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """[a | b] along the last axis for two validated (B, n) uint8 arrays.  When both row lengths are multiples of eight
    the rows are copied as 64-bit words (numpy's strided byte copy runs at 1.5 GB/s, this at memory bandwidth)."""
    if a.ndim == 2 and b.ndim == 2 and a.shape[0] == b.shape[0] and a.shape[1] % 8 == 0 and b.shape[1] % 8 == 0 and \
            a.flags.c_contiguous and b.flags.c_contiguous and a.dtype == np.uint8 and b.dtype == np.uint8:
        na, nb = a.shape[1] // 8, b.shape[1] // 8
        out = empty_big((a.shape[0], a.shape[1] + b.shape[1]), np.uint8)
        ow = out.view(np.uint64)
        ow[:, :na] = np.asarray(a).view(np.uint64)
        ow[:, na:] = np.asarray(b).view(np.uint64)
        return out
    return np.concatenate([a, b], axis=-1)


def _has_non_bit(a: np.ndarray) -> bool:
    """Any byte above 1?  One pass at memory bandwidth (eight bytes per word) instead of a bytewise max."""
    if a.flags.c_contiguous and a.size >= 64:
        flat = a.reshape(-1)
        n8 = flat.size & ~7
        if np.bitwise_or.reduce(flat[:n8].view(np.uint64)) & np.uint64(0xFEFEFEFEFEFEFEFE):
            return True
        return bool(n8 < flat.size and flat[n8:].max() > 1)
    return bool(a.max() > 1)


def words_for(nbits: int) -> int:
    return max(1, (nbits + 63) // 64)


def pack_bits(bits: np.ndarray) -> np.ndarray:
    """(B, nbits) uint8 0/1  ->  (B, ceil(nbits/64)) uint64, bit k of word w = bit 64w+k (compresscol layout)."""
    bits = np.asarray(bits, dtype=np.uint8)
    if bits.ndim == 1:
        bits = bits[None, :]
    B, n = bits.shape
    W = words_for(n)
    pad = np.zeros((B, W * 64), dtype=np.uint8)
    pad[:, :n] = bits
    by = np.packbits(pad, axis=1, bitorder="little")          # (B, W*8) bytes, little-endian bit order
    return np.ascontiguousarray(by).view("<u8").reshape(B, W)


def unpack_bits(words: np.ndarray, nbits: int) -> np.ndarray:
    """Inverse of `pack_bits`."""
    words = np.ascontiguousarray(words, dtype="<u8")
    if words.ndim == 1:
        words = words[None, :]
    B = words.shape[0]
    if B == 0:
        return np.zeros((0, nbits), dtype=np.uint8)
    by = words.view(np.uint8).reshape(B, -1)
    return np.unpackbits(by, axis=1, bitorder="little")[:, :nbits].copy()


def pack_rows(M: np.ndarray) -> np.ndarray:
    """Pack each ROW of a 0/1 matrix (rows x nbits) into uint64 words (rows, ceil(nbits/64))."""
    return pack_bits(np.asarray(M, dtype=np.uint8))


def bitmul(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """GF(2) matrix product through the packed popcount-parity form of `bitmul!` (mod2.jl:44-57).

    C[i, j] = parity( sum_k popcount(ca[k, i] & cb[k, j]) ), with ca = compresscol(A'), cb = compresscol(B).
    """
    A = as_bits(A)
    B = as_bits(B)
    ca = pack_rows(A)                 # (m, W): row i of A packed over its columns   == compresscol(A')[:, i]
    cb = pack_rows(B.T)               # (n, W): column j of B packed over its rows   == compresscol(B)[:, j]
    x = ca[:, None, :] & cb[None, :, :]
    # popcount parity of a uint64: fold by XOR
    for s in (32, 16, 8, 4, 2, 1):
        x = x ^ (x >> np.uint64(s))
    return (np.bitwise_xor.reduce(x, axis=2) & np.uint64(1)).astype(np.uint8)
