"""Lookup-table decoder, batched on the GPU: a baseline that shares the GF(2) front and back end of the tensor-network
decoders.

Reference: src/decoding/truthtable.jl -- `TruthTable` (:12-16), `make_table` (:51-90: every error of weight <= d, weights
in increasing order, qubit subsets in lexicographic order, per-qubit Paulis X, Y, Z in that order; a syndrome keeps its
first pattern unless a later one is strictly more probable, :83-87, 168-173), `TableDecoder` / `compile` / `decode`
(:196-209, 121-136), `save_table` / `load_table` (:138-166).

Here the enumeration is vectorised with numpy, the syndromes of all enumerated patterns come from the packed GF(2)
mat-vec (`mod2` on the host: compile time), the table lives on the device as sorted keys + values, and `decode` is one
batched binary-search kernel (`k_table_lookup` through `tqec_table_decode`).

Difference from the reference, documented: `get_probability` there takes `length(cep[1])` of an INTEGER as the qubit
count (truthtable.jl:178), which is 1, so it compares the first qubit's factor only; here the probability of the whole
pattern is used (the documented intent: "DistributionError").
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _cabi
from .error_model import CSSErrorPattern, CSSSyndrome, IndependentDepolarizingError
from .mod2 import as_bits, pack_bits, unpack_bits
from .tanner import CSSTannerGraph


@dataclass
class TruthTable:
    """Sorted syndrome keys -> (x error | z error) words.  keys: (n, ceil(num_st/64)) uint64; values: (n, ceil(2 nq/64))
    uint64 holding the 2 nq bits [x errors, z errors]."""
    keys: np.ndarray
    values: np.ndarray
    num_qubits: int
    num_st: int

    def __len__(self):
        return self.keys.shape[0]


def _sort_keys(keys: np.ndarray):
    """Row order of multi-word keys, most significant word last."""
    return np.lexsort(tuple(keys[:, w] for w in range(keys.shape[1])))


def make_table(tanner: CSSTannerGraph, d: int, pvec: Optional[IndependentDepolarizingError] = None) -> TruthTable:
    """truthtable.jl:51-90.  `pvec` = None is the reference's `UniformError` (a syndrome keeps its first pattern)."""
    n = tanner.stgx.nq
    Hx, Hz = tanner.stgx.H.astype(np.uint8), tanner.stgz.H.astype(np.uint8)
    nsx, nsz = Hx.shape[0], Hz.shape[0]
    pats_x, pats_z = [np.zeros((1, n), dtype=np.uint8)], [np.zeros((1, n), dtype=np.uint8)]
    for k in range(1, d + 1):
        combos = np.array(list(itertools.combinations(range(n), k)), dtype=np.int64).reshape(-1, k)
        digits = (np.arange(3 ** k)[:, None] // (3 ** np.arange(k))[None, :]) % 3          # digit j of pattern i
        epx = (digits <= 1).astype(np.uint8)                                              # 0 -> X, 1 -> Y, 2 -> Z
        epz = (digits >= 1).astype(np.uint8)
        nc, npat = combos.shape[0], digits.shape[0]
        X = np.zeros((nc, npat, n), dtype=np.uint8)
        Z = np.zeros((nc, npat, n), dtype=np.uint8)
        ci = np.arange(nc)[:, None, None]
        pi = np.arange(npat)[None, :, None]
        X[ci, pi, combos[:, None, :]] = epx[None, :, :]
        Z[ci, pi, combos[:, None, :]] = epz[None, :, :]
        pats_x.append(X.reshape(-1, n))
        pats_z.append(Z.reshape(-1, n))
    X, Z = np.concatenate(pats_x), np.concatenate(pats_z)
    sx = (Z.astype(np.int64) @ Hx.T.astype(np.int64)) & 1                               # X checks see Z errors
    sz = (X.astype(np.int64) @ Hz.T.astype(np.int64)) & 1
    keys = pack_bits(np.concatenate([sx, sz], axis=1).astype(np.uint8))
    prob = np.ones(X.shape[0])
    if pvec is not None:
        T = np.stack([np.stack([1 - pvec.px - pvec.py - pvec.pz, pvec.pz], axis=1),
                      np.stack([pvec.px, pvec.py], axis=1)], axis=1)                       # T[q, x, z]
        for q in range(n):                                       # the reference's product, qubit by qubit (:176-187)
            prob = prob * T[q, X[:, q], Z[:, q]]
    # first pattern per syndrome unless a later one is strictly more probable: stable sort by (key, -prob, enumeration order)
    order = np.lexsort((np.arange(X.shape[0]), -prob) + tuple(keys[:, w] for w in range(keys.shape[1])))
    ks = keys[order]
    first = np.ones(len(order), dtype=bool)
    first[1:] = (ks[1:] != ks[:-1]).any(axis=1)
    sel = order[first]
    vals = pack_bits(np.concatenate([X[sel], Z[sel]], axis=1))
    return TruthTable(np.ascontiguousarray(keys[sel]), vals, n, nsx + nsz)


@dataclass
class TableDecoder:
    """truthtable.jl:196-203: `d` = the largest error weight enumerated."""
    d: int
    device: int = 0


class CompiledTable:
    def __init__(self, table: TruthTable, tanner: CSSTannerGraph, device: int = 0):
        self.table = table
        self.tanner = tanner
        self.handle = _cabi.Table(table.keys, table.values, table.num_st, 2 * table.num_qubits, device)


def compile_table(decoder: TableDecoder, problem) -> CompiledTable:
    """truthtable.jl:205-208."""
    return CompiledTable(make_table(problem.tanner, decoder.d, problem.pvec), problem.tanner, decoder.device)


def decode_table(ct: CompiledTable, syndrome: CSSSyndrome):
    """truthtable.jl:121-136, batched: success_tag False and a zero pattern for syndromes outside the table."""
    from .decoding import DecodingResult
    single = as_bits(syndrome.sx).ndim == 1
    sx, sz = np.atleast_2d(as_bits(syndrome.sx)), np.atleast_2d(as_bits(syndrome.sz))
    if sx.shape[1] != ct.tanner.stgx.ns or sz.shape[1] != ct.tanner.stgz.ns:
        raise ValueError("syndrome size does not match the code")
    corr, found = ct.handle.decode(pack_bits(np.concatenate([sx, sz], axis=1)))
    n = ct.table.num_qubits
    e = unpack_bits(corr, 2 * n)
    if single:
        return DecodingResult(bool(found[0]), CSSErrorPattern(e[0, :n], e[0, n:]))
    return DecodingResult(found, CSSErrorPattern(e[:, :n], e[:, n:]))


def save_table(tb: TruthTable, filename: str) -> None:
    """truthtable.jl:138-146: one line per entry, tab-separated unsigned words: syndrome words, x-error words, z-error
    words (`writedlm` of the LongLongUInt contents)."""
    n = tb.num_qubits
    e = unpack_bits(tb.values, 2 * n)
    xw, zw = pack_bits(e[:, :n]), pack_bits(e[:, n:])
    with open(filename, "w") as fh:
        for k, x, z in zip(tb.keys, xw, zw):
            fh.write("\t".join(str(int(w)) for w in list(k) + list(x) + list(z)) + "\n")


def load_table(filename: str, num_qubits: int, num_st: int) -> TruthTable:
    """truthtable.jl:148-166."""
    cs, c = max(1, (num_st + 63) // 64), max(1, (num_qubits + 63) // 64)
    rows = [[int(t) for t in ln.split()] for ln in open(filename) if ln.strip()]
    data = np.array(rows, dtype=np.uint64).reshape(-1, cs + 2 * c)
    keys = data[:, :cs]
    x = unpack_bits(data[:, cs:cs + c], num_qubits)
    z = unpack_bits(data[:, cs + c:], num_qubits)
    order = _sort_keys(keys)
    return TruthTable(np.ascontiguousarray(keys[order]), pack_bits(np.concatenate([x, z], axis=1))[order], num_qubits, num_st)
