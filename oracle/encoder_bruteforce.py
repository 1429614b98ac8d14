"""Brute-force statement of the reference's encoder-network inference (oracle; see oracle/__init__.py).

`syndrome_inference` (src/decoding/inferenceswithencoder.jl:56-66) contracts the Clifford network of
src/nonclifford/tensornetwork.jl:51-117: gate tensors R[out, in] = tr(P_out U P_in U^dag) / 2^k (signed permutation
matrices, src/clifford/paulibasis.jl:95-101), prior vectors on the physical ends, projectors on the mapped ends of the
measured qubits, marginals on every mapped end.  Because every gate tensor has exactly one non-zero entry per input, the
contraction is a sum over the 4^n Pauli strings E at the mapped end:
    weight(E) = prod_gates R[out, in] * prod_i p_i[(U E U^dag)_i]       (the product of R entries is the sign +-1)
and the marginal of qubit k is the normalised sum of weight(E) over the E consistent with the syndrome, grouped by E_k.
Here: direct enumeration, gates applied to Pauli indices with 4x4 / 16x16 matrices built from the gate unitaries."""
import numpy as np

_P = [np.eye(2), np.array([[0, 1], [1, 0]]), np.array([[0, -1j], [1j, 0]]), np.array([[1, 0], [0, -1]])]
_U1 = {"H": np.array([[1, 1], [1, -1]]) / np.sqrt(2), "X": _P[1], "Y": _P[2], "Z": _P[3], "S": np.diag([1, 1j])}


def pauli_repr(U):
    """paulibasis.jl:95-101, little-endian Pauli basis (first qubit least significant)."""
    k = int(np.log2(U.shape[0]))

    def pm(idx):
        m = np.eye(1)
        for q in range(k):
            m = np.kron(_P[(idx >> (2 * q)) & 3], m)
        return m
    N = 4 ** k
    return np.round(np.array([[np.trace(pm(i) @ U @ pm(j) @ U.conj().T).real / 2 ** k for j in range(N)] for i in range(N)]), 12)


def gate_unitary(name):
    if name == "CNOT":                       # local qubit 0 = control, 1 = target; basis index = q0 + 2 q1
        U = np.zeros((4, 4))
        for q1 in range(2):
            for q0 in range(2):
                U[(q0 ^ q1) * 2 + q0, q1 * 2 + q0] = 1
        return U
    return _U1[name]


def marginals(qc, n, p, syn):
    """qc: gates ("H", q) / ("CNOT", c, t) ...; p[i] = (pI, pX, pY, pZ); syn = {qubit: bit}.  -> {qubit: vector}: (I, Z) /
    (X, Y) for measured qubits, (I, X, Y, Z) otherwise, normalised."""
    R = {g[0]: pauli_repr(gate_unitary(g[0])) for g in qc}
    E = np.arange(4 ** n)
    cur = [(E >> (2 * q)) & 3 for q in range(n)]                 # Pauli of every qubit at the mapped end
    sign = np.ones(E.shape[0])
    for g in qc:
        qs = g[1:]
        j = sum(cur[q] << (2 * t) for t, q in enumerate(qs))
        Rg = R[g[0]]
        i = np.abs(Rg).argmax(axis=0)[j]                         # the one output Pauli per input
        sign = sign * Rg[i, j]
        for t, q in enumerate(qs):
            cur[q] = (i >> (2 * t)) & 3
    w = sign.copy()
    for q in range(n):
        w = w * np.asarray(p[q], dtype=np.float64)[cur[q]]
    ok = np.ones(E.shape[0], dtype=bool)
    for q, b in syn.items():
        e = (E >> (2 * q)) & 3
        ok &= np.isin(e, (1, 2)) if b else np.isin(e, (0, 3))
    out = {}
    for k in range(n):
        e = (E >> (2 * k)) & 3
        full = np.array([w[ok & (e == a)].sum() for a in range(4)])
        v = full[[1, 2]] if (k in syn and syn[k]) else (full[[0, 3]] if k in syn else full)
        out[k] = v / v.sum()
    return out
