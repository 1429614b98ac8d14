"""Tanner graphs and logical operators (host-side data model of the hot path).

Reference: src/codes/ldpc.jl:27-58 (`SimpleTannerGraph`), :72-115 (`CSSTannerGraph`), :76-79 (`nq`, `ns`);
src/codes/code_distance.jl:24-100 (`row_echelon_form`, `null_space`, `logical_operator`, `same_qubit_order`);
src/codes/gaussian_elimination.jl:32-96 (`gaussian_elimination!` with / without column operations).

Indices are 0-BASED here (the reference is 1-based Julia); orderings are otherwise identical, so
`s2q`, `q2s`, `H`, `lx`, `lz` equal the reference's after the -1 shift.  The representative choice of
`lx` / `lz` matters (it labels the axes of the TNMMAP marginal), so the elimination below follows
the reference's pivoting rule step by step.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

from .mod2 import as_bits


class AbstractTannerGraph:
    pass


@dataclass
class SimpleTannerGraph(AbstractTannerGraph):
    """ldpc.jl:27-33.  `H` is (ns, nq) uint8."""
    nq: int
    ns: int
    q2s: List[List[int]]
    s2q: List[List[int]]
    H: np.ndarray

    def __init__(self, nq=None, sts=None, *, H=None):
        # SimpleTannerGraph(nq, sts) (ldpc.jl:44-52)  or  SimpleTannerGraph(H) (ldpc.jl:53-58)
        if sts is None and H is None and not isinstance(nq, (int, np.integer)):
            H, nq = nq, None
        if H is not None:
            H = as_bits(H)
            nq = H.shape[1]
            sts = [list(map(int, np.flatnonzero(H[i]))) for i in range(H.shape[0])]
        sts = [list(map(int, s)) for s in sts]
        for s in sts:
            for q in s:
                if not 0 <= q < nq:
                    raise IndexError(f"check touches bit {q} outside 0..{nq - 1}")   # BoundsError in the reference
        self.nq = int(nq)
        self.ns = len(sts)
        self.s2q = sts
        self.q2s = [[i for i, s in enumerate(sts) if q in s] for q in range(self.nq)]
        self.H = np.zeros((self.ns, self.nq), dtype=np.uint8)
        for i, s in enumerate(sts):
            for q in s:
                self.H[i, q] = 1


@dataclass
class CSSTannerGraph(AbstractTannerGraph):
    """ldpc.jl:72-75.  `stgx`: X stabilizers (detect Z errors); `stgz`: Z stabilizers (detect X errors)."""
    stgx: SimpleTannerGraph
    stgz: SimpleTannerGraph

    def __init__(self, *args):
        if len(args) == 1:
            a = args[0]
            if hasattr(a, "stabilizers"):            # CSSTannerGraph(code::CSSQuantumCode), ldpc.jl:113-115
                a = a.stabilizers()
            # CSSTannerGraph(sts::Vector{PauliString}) (ldpc.jl:105-111): a stabilizer is X-type / Z-type by its
            # first non-identity Pauli; here a stabilizer is a pair (kind, support) with kind in "XZ".
            n = a.nq
            stxs = [list(sup) for kind, sup in a.rows if kind == "X"]
            stzs = [list(sup) for kind, sup in a.rows if kind == "Z"]
            self.stgx, self.stgz = SimpleTannerGraph(n, stxs), SimpleTannerGraph(n, stzs)
        elif len(args) == 2:
            self.stgx, self.stgz = args
        elif len(args) == 3:                          # CSSTannerGraph(nq, stxs, stzs), ldpc.jl:101-103
            n, stxs, stzs = args
            self.stgx, self.stgz = SimpleTannerGraph(n, stxs), SimpleTannerGraph(n, stzs)
        else:
            raise TypeError("CSSTannerGraph(code) | (stgx, stgz) | (nq, stxs, stzs)")
        if self.stgx.nq != self.stgz.nq:
            raise ValueError("X and Z Tanner graphs must have the same number of qubits")


def nq(t: AbstractTannerGraph) -> int:
    return t.nq if isinstance(t, SimpleTannerGraph) else t.stgx.nq


def ns(t: AbstractTannerGraph) -> int:
    return t.ns if isinstance(t, SimpleTannerGraph) else t.stgx.ns + t.stgz.ns


@dataclass
class StabilizerList:
    """A list of CSS stabilizer generators: rows = [("X" | "Z", support tuple)], in generation order."""
    nq: int
    rows: list = field(default_factory=list)

    def __len__(self):
        return len(self.rows)


# ------------------------------------------------------------------------------------------------------------
# GF(2) elimination, restating gaussian_elimination.jl:32-96 for the `SimpleBimatrix` case (offsets = 0).
# ------------------------------------------------------------------------------------------------------------
def _eliminate_with_column_swaps(M: np.ndarray):
    """Row by row: bring the first non-zero of row i to column (#pivots so far) by a COLUMN swap, then clear that
    column from every other row.  Returns (matrix, ordering) with ordering[c] = original column now at c."""
    M = M.copy().astype(np.uint8)
    nr, nc = M.shape
    ordering = list(range(nc))
    zero_rows = 0
    for i in range(nr):
        nzc = np.flatnonzero(M[i])
        if nzc.size == 0:
            zero_rows += 1
            continue
        tgt = i - zero_rows
        j = int(nzc[0])
        if j != tgt:
            M[:, [tgt, j]] = M[:, [j, tgt]]
            ordering[tgt], ordering[j] = ordering[j], ordering[tgt]
        hit = np.flatnonzero(M[:, tgt])
        for k in hit:
            if k != i:
                M[k] ^= M[i]
    return M, ordering


def row_echelon_form(H: np.ndarray) -> np.ndarray:
    """code_distance.jl:24-28: reduced row echelon form by ROW swaps only (gaussian_elimination.jl:51-81)."""
    M = as_bits(H).copy()
    nr, nc = M.shape
    zero_col = 0
    for i in range(nr):
        if i + zero_col >= nc:
            return M
        col = i + zero_col
        nz = np.flatnonzero(M[i:, col])
        while nz.size == 0:
            zero_col += 1
            if i + zero_col >= nc:
                return M
            col = i + zero_col
            nz = np.flatnonzero(M[i:, col])
        j = i + int(nz[0])
        if j != i:
            M[[i, j]] = M[[j, i]]
        for k in np.flatnonzero(M[:, col]):
            if k != i:
                M[k] ^= M[i]
    return M


def null_space(H: np.ndarray) -> np.ndarray:
    """code_distance.jl:30-51: one basis vector per free column j: e_j + sum_{i: reH[i,j]} e_{pivot_i}."""
    re = row_echelon_form(H)
    m, n = re.shape
    pivots = [int(np.flatnonzero(re[i])[0]) for i in range(m) if re[i].any()]
    out = np.zeros((n - len(pivots), n), dtype=np.uint8)
    r = 0
    pset = set(pivots)
    for j in range(n):
        if j in pset:
            continue
        out[r, j] = 1
        for i, p in enumerate(pivots):
            if re[i, j]:
                out[r, p] = 1
        r += 1
    return out


def _logical_operator_one(Hx: np.ndarray, Hz: np.ndarray) -> np.ndarray:
    """code_distance.jl:53-64: rows of ker(Hx) that survive elimination against Hz."""
    ker = null_space(Hx)
    H = np.vstack([as_bits(Hz), ker]).astype(np.uint8)
    M, ordering = _eliminate_with_column_swaps(H)
    lz = M[Hz.shape[0]:]
    lz = lz[lz.any(axis=1)]
    out = np.zeros_like(lz)
    out[:, ordering] = lz                      # lz[:, bimat.ordering] = lz
    return out


def same_qubit_order(lx: np.ndarray, lz: np.ndarray):
    """code_distance.jl:84-100: pair lx[j] with lz[j] so that lx[i].lz[j] = delta_ij."""
    lxc = lx.copy()
    lx_new = np.zeros_like(lx)
    for j in range(lz.shape[0]):
        hits = [i for i in range(lxc.shape[0]) if int(lxc[i] @ lz[j]) & 1]
        if not hits:
            raise AssertionError("The logical operator is linearly dependent!")
        lx_new[j] = lxc[hits[0]]
        for i in hits[1:]:
            lxc[i] ^= lxc[hits[0]]
    return lx_new, lz


def logical_operator(tanner: CSSTannerGraph):
    """code_distance.jl:78-82 -> (lx, lz), each (k, n) uint8."""
    lz = _logical_operator_one(tanner.stgx.H, tanner.stgz.H)
    lx = _logical_operator_one(tanner.stgz.H, tanner.stgx.H)
    return same_qubit_order(lx, lz)


def gf2_right_inverse(H: np.ndarray):
    """Particular-solution map for H e = s over GF(2): returns (R, rank) with R (nq, ns) such that
    H (R s) = s for every s in the column space of H.  Replaces the per-shot SCIP feasibility program
    `_mixed_integer_programming_for_one_solution` (src/decoding/ipdecoder.jl:150-169): the representative differs
    from SCIP's by an element of ker(H) (documented gauge difference)."""
    H = as_bits(H)
    m, n = H.shape
    A = np.concatenate([H.copy(), np.eye(m, dtype=np.uint8)], axis=1)     # [H | I]: row ops tracked on the right
    piv_cols = []
    r = 0
    for c in range(n):
        nz = np.flatnonzero(A[r:, c]) if r < m else np.zeros(0, dtype=int)
        if nz.size == 0:
            continue
        p = r + int(nz[0])
        if p != r:
            A[[r, p]] = A[[p, r]]
        for k in np.flatnonzero(A[:, c]):
            if k != r:
                A[k] ^= A[r]
        piv_cols.append(c)
        r += 1
        if r == m:
            break
    T = A[:, n:]                                   # T H = rref(H);  e[piv_cols[i]] = (T s)[i], free vars = 0
    R = np.zeros((n, m), dtype=np.uint8)
    for i, c in enumerate(piv_cols):
        R[c] = T[i]
    return R, len(piv_cols)


def gf2_sector_fixes(H: np.ndarray, L: np.ndarray) -> np.ndarray:
    """Undetectable patterns that move between logical sectors: rows f_j with H f_j = 0 whose sector flips
    d_j = L f_j form a basis, in REDUCED ECHELON form (the lowest set bit of d_j is its pivot; no other row has that
    bit), of the reachable flips {L f : H f = 0}.  `tqec_coset_rep` walks these rows once to move the representative
    R s into the decoded sector -- also when only a joint flip of several observables is undetectable, which a
    per-observable solve (H f = 0, L f = e_l) cannot express.  -> (r, n) uint8, r <= rows(L)."""
    H, L = as_bits(H), as_bits(L)
    n = H.shape[1]
    k = L.shape[0]
    if k == 0:
        return np.zeros((0, n), dtype=np.uint8)
    # null space of H by elimination on [H^T | I]: rows whose H^T part vanishes carry a kernel vector on the right
    A = np.concatenate([H.T.copy(), np.eye(n, dtype=np.uint8)], axis=1)
    m = H.shape[0]
    r = 0
    for c in range(m):
        nz = np.flatnonzero(A[r:, c]) if r < n else np.zeros(0, dtype=int)
        if nz.size == 0:
            continue
        p = r + int(nz[0])
        if p != r:
            A[[r, p]] = A[[p, r]]
        rows = np.flatnonzero(A[:, c])
        rows = rows[rows != r]
        A[rows] ^= A[r]
        r += 1
        if r == n:
            break
    ker = A[r:, m:]                                             # (n - rank, n), H ker^T = 0
    D = ((ker.astype(np.int64) @ L.T.astype(np.int64)) & 1).astype(np.uint8)    # sector flip of every kernel vector
    M = np.concatenate([D, ker], axis=1)                        # reduced row echelon form on the sector part
    r = 0
    for j in range(k):
        nz = np.flatnonzero(M[r:, j]) if r < M.shape[0] else np.zeros(0, dtype=int)
        if nz.size == 0:
            continue
        p = r + int(nz[0])
        if p != r:
            M[[r, p]] = M[[p, r]]
        rows = np.flatnonzero(M[:, j])
        rows = rows[rows != r]
        M[rows] ^= M[r]
        r += 1
    return np.ascontiguousarray(M[:r, k:])
