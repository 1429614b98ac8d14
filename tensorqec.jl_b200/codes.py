"""Stabilizer generators of the codes named by the benchmark configs (input fixtures of the hot path).

Reference: src/codes/codes.jl:9-66 (`SurfaceCode`), :98-110 (`SteaneCode`), :247-249 + :256-334 + :410-421
(`Color488`, all X-type strings first, then the same supports as Z-type).  Qubits are 0-based here; qubit (i, j)
of an m x n surface code (row i, column j, 0-based) has index i*n + j, which is the reference's
`reshape(1:m*n, n, m)'` numbering minus one.  Generator ORDER is the reference's, because it fixes the order of
the syndrome bits (`CSSTannerGraph(sts)` keeps generation order inside each Pauli type, src/codes/ldpc.jl:105-111).
Supports are stored ascending (the reference's `findall` over a PauliString).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .tanner import StabilizerList


class QuantumCode:
    pass


class CSSQuantumCode(QuantumCode):
    def stabilizers(self) -> StabilizerList:            # pragma: no cover - interface
        raise NotImplementedError


def stabilizers(code: CSSQuantumCode) -> StabilizerList:
    return code.stabilizers()


@dataclass(frozen=True)
class SurfaceCode(CSSQuantumCode):
    """Rotated surface code with m rows and n columns (codes.jl:9-12, 37-66)."""
    m: int
    n: int

    def stabilizers(self) -> StabilizerList:
        m, n = self.m, self.n
        q = lambda i, j: (i - 1) * n + (j - 1)          # (i, j) 1-based as in the reference -> 0-based qubit
        rows = []

        def plaquettes(kind, parity):
            for i in range(1, m):
                for j in range(1, n):
                    if (i + j) % 2 == parity:
                        rows.append((kind, tuple(sorted((q(i, j), q(i + 1, j), q(i, j + 1), q(i + 1, j + 1))))))

        plaquettes("X", 0)
        o = (n + 1) % 2
        for i in range(1, m // 2 + 1):                  # X boundary: right column, then left column
            if 2 * i + o <= m:
                rows.append(("X", tuple(sorted((q(2 * i - 1 + o, n), q(2 * i + o, n))))))
            if 2 * i + 1 <= m:
                rows.append(("X", tuple(sorted((q(2 * i, 1), q(2 * i + 1, 1))))))
        plaquettes("Z", 1)
        e = m % 2
        for j in range(1, n // 2 + 1):                  # Z boundary: top row, then bottom row
            rows.append(("Z", tuple(sorted((q(1, 2 * j - 1), q(1, 2 * j))))))
            if 2 * j + e <= n:
                rows.append(("Z", tuple(sorted((q(m, 2 * j - 1 + e), q(m, 2 * j + e))))))
        return StabilizerList(m * n, rows)


@dataclass(frozen=True)
class SteaneCode(CSSQuantumCode):
    """[[7,1,3]] (codes.jl:98-110)."""

    def stabilizers(self) -> StabilizerList:
        sup = [(0, 2, 4, 6), (1, 2, 5, 6), (3, 4, 5, 6)]
        return StabilizerList(7, [("X", s) for s in sup] + [("Z", s) for s in sup])


@dataclass(frozen=True)
class Color488(CSSQuantumCode):
    """Triangular 4.8.8 colour code of odd distance d (codes.jl:247-249; check matrix :256-334).

    The lattice is built layer by layer (layer L adds 4L qubits): red squares, one half-octagon on the left
    (odd L) or right (even L) boundary, L half-octagons along the bottom which the next layer completes into
    full octagons.  d = 5 gives the 8 x 17 matrix with row weights [4,4,8,4,4,4,4,4] listed in SURVEY C.4.
    """
    d: int

    def check_matrix(self) -> np.ndarray:
        d = self.d
        if d < 3 or d % 2 == 0:
            raise ValueError("Color488 needs an odd distance >= 3")
        n = (d * d + 2 * d - 1) // 2
        layers = (d - 1) // 2
        H = np.zeros(((n - 1) // 2, n), dtype=np.uint8)
        row = 0
        base = 0                                           # 0-based index of the first qubit of this layer
        last_even = layers % 2 == 0
        for L in range(1, layers + 1):
            last = L == layers
            shift = 1 if (last and last_even) else 0
            step = 2 * L
            # the L-1 bottom half-octagons of the previous layer become full octagons
            row -= L - 1
            for j in range(L - 1):
                a = base + step + 1 + 2 * j
                H[row, [a, a + 1, a + step + 1 - shift, a + step + 2 - shift]] = 1
                row += 1
            for j in range(L):                              # red squares
                a = base + 2 * j
                H[row, [a, a + 1, a + step, a + step + 1]] = 1
                row += 1
            if L % 2 == 1:                                  # half-octagon on the left boundary
                H[row, [base, base + step, base + 2 * step, base + 2 * step + 1]] = 1
            else:                                           # half-octagon on the right boundary
                a = base + 1 + 2 * (L - 1)
                off = 1 if last else 0
                H[row, [a, a + step, base + 3 * step - off, base + 3 * step + 1 - off]] = 1
            row += 1
            for j in range(L):                              # half-octagons along the bottom
                a = base + step + 2 * j
                H[row, [a, a + 1, a + step + 1 - shift, a + step + 2 - shift]] = 1
                row += 1
            base += 4 * L
        return H

    def stabilizers(self) -> StabilizerList:
        H = self.check_matrix()
        sup = [tuple(int(c) for c in np.flatnonzero(r)) for r in H]
        return StabilizerList(H.shape[1], [("X", s) for s in sup] + [("Z", s) for s in sup])
