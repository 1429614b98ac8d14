"""Encoder-circuit ("Ferris-Poulin") inference, batched on the GPU.

Reference: src/codes/encoder.jl (`CSSBimatrix` :16-21, `stabilizers2bimatrix` :41-48, `encode_circuit` :57-78,
`encode_stabilizers` :94-100), src/codes/gaussian_elimination.jl (`switch_qubits!` :1-9, `gaussian_elimination!`
:32-103), src/nonclifford/tensornetwork.jl (`clifford_network` :51-62, `generate_tensor_network` :72-117),
src/decoding/inferenceswithencoder.jl (`syndrome_transform` :18-20, `generate_syndrome_dict` :22-24,
`syndrome_inference` :56-66, `correction_pauli_string` :81-96, `inference` :112-118), src/clifford/paulibasis.jl
(`pauli_repr` :95-101, `pauli_string_map_iter` :113-135).

What the reference contracts.  The encoding circuit U (H and CNOT gates) becomes a tensor network over labels of
dimension 4 (one Pauli per wire segment): every gate contributes its Pauli-representation tensor
R[out, in] = tr(P_out U P_in U^dag) / 2^k -- a SIGNED permutation matrix: U P_in U^dag = +-P_out -- the prior vector
p_i = (p_I, p_X, p_Y, p_Z) sits on the physical (output) end of wire i, the mapped (input) end of a measured qubit
carries the projector on {I, Z} (syndrome bit 0) or {X, Y} (bit 1), and `TensorInference.marginals` returns, for every
qubit, the normalised marginal of the Pauli at its mapped end.  In words: m_k[P] = sum over mapped Pauli strings E with
E_k = P that agree with the syndrome of  sign(E) * prod_i p_i[(U E U^dag)_i].  (The sign is part of the reference's
tensors, so it is part of this restatement; it equals +1 for every string when p_Y terms never meet an H or a Y-type
CNOT input.)

How it runs here.  A label of dimension 4 is two bits (x, z) with I = (0,0), X = (1,0), Y = (1,1), Z = (0,1); a Clifford
gate maps the bits LINEARLY over GF(2), so a gate tensor is a set of parity constraints "output bit = xor of input bits"
plus a +-1 table over its input bits.  That is exactly the input format of the sum-product frontier executor
(`tqec_plan_compile`): factors = gate sign tables and priors, rows = the gates' bit relations (clamped to a constant 0),
the syndrome clamps of the measured qubits, and the two bits of the queried qubit as open axes.  One plan per qubit,
every plan batched over syndromes.  A 2n-bit frontier (18 bits for the 9-qubit surface code) runs on the global-memory
executor, smaller ones on chip.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _cabi, schedule as S
from .mod2 import pack_bits
from .tanner import CSSTannerGraph

PAULI_BITS = [(0, 0), (1, 0), (1, 1), (0, 1)]                    # I, X, Y, Z -> (x, z)
BITS_PAULI = {(0, 0): 0, (1, 0): 1, (1, 1): 2, (0, 1): 3}


# ---- encoding circuit (encoder.jl, gaussian_elimination.jl) ---------------------------------------------------------
@dataclass
class CSSBimatrix:
    matrix: np.ndarray            # (n_stabilizers, 2n) uint8: [X part | Z part]
    Q: np.ndarray                 # (n_stabilizers, n_stabilizers) uint8: row operations of the elimination
    ordering: List[int]           # 0-based qubit order after the column swaps
    xcodenum: int


def stabilizers2bimatrix(tanner: CSSTannerGraph) -> CSSBimatrix:
    """encoder.jl:41-48: X-type generators first (acting on the X half), then Z-type generators (Z half)."""
    A, B = tanner.stgx.H.astype(np.uint8), tanner.stgz.H.astype(np.uint8)
    n = A.shape[1]
    M = np.zeros((A.shape[0] + B.shape[0], 2 * n), dtype=np.uint8)
    M[: A.shape[0], :n] = A
    M[A.shape[0]:, n:] = B
    return CSSBimatrix(M, np.eye(M.shape[0], dtype=np.uint8), list(range(n)), A.shape[0])


def _switch_qubits(b: CSSBimatrix, i: int, j: int):
    n = b.matrix.shape[1] // 2
    b.ordering[i], b.ordering[j] = b.ordering[j], b.ordering[i]
    b.matrix[:, [i, j]] = b.matrix[:, [j, i]]
    b.matrix[:, [n + i, n + j]] = b.matrix[:, [n + j, n + i]]


def _eliminate(b: CSSBimatrix, rows: range, col_offset: int, qubit_offset: int):
    """gaussian_elimination.jl:32-103 with column operations allowed (0-based restatement)."""
    start_col = col_offset + qubit_offset
    zero_row = 0
    for i in rows:
        offset = i - rows.start - zero_row
        nz = np.flatnonzero(b.matrix[i, start_col:])
        if nz.size == 0:
            zero_row += 1
            continue
        _switch_qubits(b, qubit_offset + offset, int(nz[0]) + qubit_offset)
        for k in rows:
            if k != i and b.matrix[k, offset + start_col]:
                b.matrix[k] ^= b.matrix[i]
                b.Q[k] ^= b.Q[i]


def gaussian_elimination(b: CSSBimatrix) -> CSSBimatrix:
    """gaussian_elimination.jl:104-109."""
    n = b.matrix.shape[1] // 2
    _eliminate(b, range(0, b.xcodenum), 0, 0)
    _eliminate(b, range(b.xcodenum, b.matrix.shape[0]), n, b.xcodenum)
    return b


Gate = Tuple                                                  # ("H", q) | ("CNOT", control, target) | ("X"|"Y"|"Z"|"S", q)


def encode_circuit(b: CSSBimatrix) -> List[Gate]:
    """encoder.jl:57-78."""
    n = b.matrix.shape[1] // 2
    rows = b.matrix.shape[0]
    qc: List[Gate] = [("H", b.ordering[i]) for i in range(b.xcodenum)]
    for i in range(b.xcodenum, rows):
        for j in range(n + rows, 2 * n):
            if b.matrix[i, j]:
                qc.append(("CNOT", b.ordering[j - n], b.ordering[i]))
    for i in range(b.xcodenum):
        for j in range(b.xcodenum, n):
            if b.matrix[i, j]:
                qc.append(("CNOT", b.ordering[i], b.ordering[j]))
    return qc


def encode_stabilizers(tanner: CSSTannerGraph):
    """encoder.jl:94-100 -> (circuit, data qubits, bimatrix)."""
    b = gaussian_elimination(stabilizers2bimatrix(tanner))
    return encode_circuit(b), b.ordering[b.matrix.shape[0]:], b


def syndrome_transform(b: CSSBimatrix, measure_outcome: np.ndarray) -> np.ndarray:
    """inferenceswithencoder.jl:18-20: Q (syn == -1); batched over the leading axis."""
    s = (np.asarray(measure_outcome) == -1).astype(np.int64)
    return ((s @ b.Q.T.astype(np.int64)) & 1).astype(np.uint8)


def generate_syndrome_dict(b: CSSBimatrix, syn: np.ndarray) -> Dict[int, np.ndarray]:
    """inferenceswithencoder.jl:22-24: transformed syndrome bit i belongs to qubit ordering[i]."""
    syn = np.asarray(syn)
    return {b.ordering[i]: syn[..., i] for i in range(b.Q.shape[1])}


# ---- Pauli representation of the gates (paulibasis.jl:95-101) --------------------------------------------------------
_P = [np.eye(2), np.array([[0, 1], [1, 0]]), np.array([[0, -1j], [1j, 0]]), np.array([[1, 0], [0, -1]])]
_U1 = {"H": np.array([[1, 1], [1, -1]]) / np.sqrt(2), "X": _P[1], "Y": _P[2], "Z": _P[3], "S": np.diag([1, 1j])}


def _pauli_mat(idx: int, k: int) -> np.ndarray:
    """k-qubit Pauli of little-endian index idx = sum_q c_q 4^q (qubit 0 = least significant, as in pauli_basis)."""
    m = np.eye(1)
    for q in range(k):
        m = np.kron(_P[(idx >> (2 * q)) & 3], m)             # qubit 0 is the rightmost factor (little endian)
    return m


def gate_action(name: str):
    """-> (k, perm, sign): U P_in U^dag = sign[in] * P_{perm[in]} for the k-qubit gate, indices little-endian."""
    if name == "CNOT":
        k = 2
        U = np.zeros((4, 4))
        for q1 in range(2):
            for q0 in range(2):                                   # local qubit 0 = control, 1 = target (convert_to_put)
                U[(q0 ^ q1) * 2 + q0, q1 * 2 + q0] = 1
    else:
        k, U = 1, _U1[name]
    N = 4 ** k
    perm, sign = np.zeros(N, dtype=np.int64), np.zeros(N)
    for j in range(N):
        img = U @ _pauli_mat(j, k) @ U.conj().T
        for i in range(N):
            c = np.trace(_pauli_mat(i, k) @ img).real / 2 ** k
            if abs(c) > 0.5:
                perm[j], sign[j] = i, np.sign(c)
    return k, perm, sign


@dataclass
class CliffordNetwork:
    """tensornetwork.jl:1-11, in bit form: `gates[g]` = (name, qubits, input labels, output labels); label l has the
    bit variables (2l, 2l + 1) = (x, z)."""
    n: int
    gates: List[Tuple[str, Tuple[int, ...], Tuple[int, ...], Tuple[int, ...]]]
    mapped: List[int]             # label at the circuit's input end of every qubit
    physical: List[int]           # label at the output end
    n_labels: int


def clifford_network(qc: Sequence[Gate], n: int) -> CliffordNetwork:
    """tensornetwork.jl:51-62."""
    pins = list(range(n))
    gates, nl = [], n
    for g in qc:
        qs = tuple(g[1:])
        ins = tuple(pins[q] for q in qs)
        outs = tuple(range(nl, nl + len(qs)))
        nl += len(qs)
        for q, o in zip(qs, outs):
            pins[q] = o
        gates.append((g[0], qs, ins, outs))
    return CliffordNetwork(n, gates, list(range(n)), pins, nl)


def _bits_of(idx: int, k: int):
    """little-endian k-qubit Pauli index -> bit vector (x_0, z_0, x_1, z_1, ...)."""
    out = []
    for q in range(k):
        out += list(PAULI_BITS[(idx >> (2 * q)) & 3])
    return out


def _idx_of(bits):
    return sum(BITS_PAULI[(bits[2 * q], bits[2 * q + 1])] << (2 * q) for q in range(len(bits) // 2))


def inference_graph(cl: CliffordNetwork, p: Sequence[Sequence[float]], measured: Sequence[int], query: int):
    """Factor graph of `syndrome_inference` for the marginal of qubit `query` -> (factors, checks, n_vars, n_checks, n_obs).
    Syndrome bit i clamps the x bit of measured[i]; the last syndrome bit is a constant 0 that closes the gates' relations."""
    factors, checks = [], []
    const0 = len(measured)
    for name, qs, ins, outs in cl.gates:
        k, perm, sign = gate_action(name)
        in_vars = [v for l in ins for v in (2 * l, 2 * l + 1)]
        out_vars = [v for l in outs for v in (2 * l, 2 * l + 1)]
        # the bit map is linear: image of every single input bit
        cols = [_bits_of(int(perm[_idx_of([int(b == a) for b in range(2 * k)])]), k) for a in range(2 * k)]
        for j in range(4 ** k):                                  # (verify linearity once per gate type: cheap)
            bits = _bits_of(j, k)
            img = [0] * (2 * k)
            for a in range(2 * k):
                if bits[a]:
                    img = [x ^ y for x, y in zip(img, cols[a])]
            assert img == _bits_of(int(perm[j]), k), "gate is not a Clifford gate"
        for r in range(2 * k):
            checks.append(S.Check(tuple([out_vars[r]] + [in_vars[a] for a in range(2 * k) if cols[a][r]]), "syn", const0))
        tab = np.array([sign[_idx_of([(a >> b) & 1 for b in range(2 * k)])] for a in range(1 << (2 * k))])
        factors.append(S.Factor(tuple(in_vars), tab))
    for q, l in enumerate(cl.physical):
        pq = np.asarray(p[q], dtype=np.float64)
        factors.append(S.Factor((2 * l, 2 * l + 1), np.array([pq[BITS_PAULI[(a & 1, a >> 1)]] for a in range(4)])))
    for i, q in enumerate(measured):
        checks.append(S.Check((2 * cl.mapped[q],), "syn", i))
    l = cl.mapped[query]
    if query in measured:
        checks.append(S.Check((2 * l + 1,), "obs", 0))
        n_obs = 1
    else:
        checks += [S.Check((2 * l,), "obs", 0), S.Check((2 * l + 1,), "obs", 1)]
        n_obs = 2
    return factors, checks, 2 * cl.n_labels, const0 + 1, n_obs


class CompiledInference:
    """`syndrome_inference` compiled for a circuit, a set of measured qubits and a prior: one sum-product plan per
    qubit; `marginals(syn)` decodes a batch of transformed syndromes (bit i = measured[i])."""

    def __init__(self, qc: Sequence[Gate], n: int, p: Sequence[Sequence[float]], measured: Sequence[int], device: int = 0):
        self.cl = clifford_network(qc, n)
        self.n, self.measured, self.p = n, list(measured), [list(x) for x in p]
        self.plans = []
        for k in range(n):
            f, c, nv, nc, no = inference_graph(self.cl, p, self.measured, k)
            self.plans.append(_cabi.Plan.compile(_cabi.Problem(f, c, S.SUMPROD, nv, nc, no, device=device)))

    def marginals(self, syn: np.ndarray) -> List[np.ndarray]:
        """syn: (B, len(measured)) 0/1.  -> per qubit k an array (B, 2) for measured qubits -- (I, Z) if its syndrome bit
        is 0, (X, Y) if 1 -- or (B, 4) in the order (I, X, Y, Z) (the reference's vectors, inferenceswithencoder.jl:52),
        each row normalised to sum 1."""
        syn = np.atleast_2d(np.asarray(syn, dtype=np.uint8))
        words = pack_bits(np.concatenate([syn, np.zeros((syn.shape[0], 1), dtype=np.uint8)], axis=1))
        out = []
        for k, plan in enumerate(self.plans):
            mar, _ = plan.decode_marginal(words)
            if k in self.measured:
                v = mar.copy()                                    # index = z bit: (I, Z) for syndrome 0; (X, Y) for syndrome 1
            else:
                v = mar[:, [0, 1, 3, 2]]                          # index x + 2 z: I, X, Z, Y -> (I, X, Y, Z)
            tot = v.sum(axis=1, keepdims=True)
            with np.errstate(invalid="ignore", divide="ignore"):
                out.append(v / tot)
        return out


def syndrome_inference(qc: Sequence[Gate], n: int, syn: Dict[int, int], p: Sequence[Sequence[float]], device: int = 0):
    """inferenceswithencoder.jl:56-66 for ONE syndrome dictionary {qubit: bit} -> {qubit: marginal vector}."""
    measured = sorted(syn)
    ci = CompiledInference(qc, n, p, measured, device)
    m = ci.marginals(np.array([[int(syn[q]) for q in measured]], dtype=np.uint8))
    return {k: m[k][0] for k in range(n)}


def correction_pauli_string(n: int, syn: Dict[int, int], prob: Dict[int, np.ndarray]) -> List[int]:
    """inferenceswithencoder.jl:81-96 -> Pauli ids (0 I, 1 X, 2 Y, 3 Z) per qubit, in the coding space."""
    ps = [0] * n
    for k, v in prob.items():
        a = int(np.argmax(v))
        if k in syn:
            if syn[k]:
                ps[k] = 1 if a == 0 else 2
            elif a == 1:
                ps[k] = 3
        else:
            ps[k] = a
    return ps


def pauli_string_map_iter(ps: Sequence[int], qc: Sequence[Gate]) -> List[int]:
    """paulibasis.jl:113-135: push a Pauli string through the circuit gate by gate (signs dropped)."""
    ps = list(ps)
    for g in qc:
        k, perm, _ = gate_action(g[0])
        qs = g[1:]
        j = sum(ps[q] << (2 * i) for i, q in enumerate(qs))
        i = int(perm[j])
        for t, q in enumerate(qs):
            ps[q] = (i >> (2 * t)) & 3
    return ps


def inference(measure_outcome: Sequence[int], code: CSSBimatrix, qc: Sequence[Gate], p: Sequence[Sequence[float]], device: int = 0):
    """inferenceswithencoder.jl:112-118 -> the physical-space correction as Pauli ids per qubit."""
    n = code.matrix.shape[1] // 2
    syn = {q: int(b) for q, b in generate_syndrome_dict(code, syndrome_transform(code, np.asarray(measure_outcome))).items()}
    pinf = syndrome_inference(qc, n, syn, p, device)
    return pauli_string_map_iter(correction_pauli_string(n, syn, pinf), qc)
