"""Encoder-circuit inference (tensorqec.jl_b200/encoder.py) on CPU: the encoding circuit against the stabilizer
formalism, the factor graph handed to the sum-product executor against brute-force enumeration of the reference's
Clifford network (oracle/encoder_bruteforce.py), through the recurrence oracle."""
import numpy as np
import pytest

import tensorqec.jl_b200 as tq
from oracle import encoder_bruteforce as bf, frontier
from tensorqec.jl_b200 import encoder as E, schedule as S


def _symplectic_images(qc, n):
    """image of every single-qubit X_q / Z_q at the circuit input, as (x bits, z bits) at the output."""
    out = {}
    for q in range(n):
        for name, pid in (("X", 1), ("Z", 3)):
            ps = [0] * n
            ps[q] = pid
            img = E.pauli_string_map_iter(ps, qc)
            out[(name, q)] = (np.array([int(a in (1, 2)) for a in img]), np.array([int(a in (2, 3)) for a in img]))
    return out


@pytest.mark.parametrize("code", ["steane", "surface3", "three"])
def test_encoding_circuit_prepares_the_stabilizer_group(code):
    """encode_stabilizers (encoder.jl:94-100): pushing Z on every ancilla (the stabilizers of |0..0> on the non-data
    qubits) through the circuit must give generators of the code's stabilizer group."""
    t = {"steane": tq.CSSTannerGraph(tq.SteaneCode()), "surface3": tq.CSSTannerGraph(tq.SurfaceCode(3, 3)),
         "three": tq.CSSTannerGraph(3, [], [[1, 2], [0, 2]])}[code]
    qc, data, b = E.encode_stabilizers(t)
    n = t.stgx.nq
    assert len(data) == n - t.stgx.ns - t.stgz.ns
    img = _symplectic_images(qc, n)
    gens = [np.concatenate(img[("Z", q)]) for q in range(n) if q not in data]
    G = np.array(gens)
    Sx = np.concatenate([t.stgx.H, np.zeros_like(t.stgx.H)], axis=1) if t.stgx.ns else np.zeros((0, 2 * n), dtype=np.uint8)
    Sz = np.concatenate([np.zeros_like(t.stgz.H), t.stgz.H], axis=1)
    Sg = np.concatenate([Sx, Sz]).astype(np.uint8)

    def rank(M):
        M = M.copy() % 2
        r = 0
        for c in range(M.shape[1]):
            nz = np.flatnonzero(M[r:, c])
            if nz.size == 0:
                continue
            M[[r, r + nz[0]]] = M[[r + nz[0], r]]
            for k in np.flatnonzero(M[:, c]):
                if k != r:
                    M[k] ^= M[r]
            r += 1
            if r == M.shape[0]:
                break
        return r
    assert rank(G.astype(np.uint8)) == len(gens) == rank(Sg)
    assert rank(np.concatenate([G, Sg]).astype(np.uint8)) == rank(Sg)        # same span


def test_gate_actions_are_signed_permutations():
    for name in ("H", "CNOT", "X", "Y", "Z", "S"):
        k, perm, sign = E.gate_action(name)
        R = bf.pauli_repr(bf.gate_unitary(name))
        assert sorted(perm.tolist()) == list(range(4 ** k)) and set(np.abs(sign)) == {1.0}
        for j in range(4 ** k):
            assert R[perm[j], j] == sign[j] and np.count_nonzero(R[:, j]) == 1
    _, perm, sign = E.gate_action("H")
    assert perm.tolist() == [0, 3, 2, 1] and sign.tolist() == [1, 1, -1, 1]     # H Y H = -Y


@pytest.mark.parametrize("code", ["three", "steane"])
def test_inference_graph_matches_bruteforce(code):
    """The factor graph of one queried qubit, evaluated by the recurrence oracle (the algorithm the CUDA kernels run),
    against direct enumeration of the reference's network -- for every qubit and several syndromes, signs included."""
    t = {"steane": tq.CSSTannerGraph(tq.SteaneCode()), "three": tq.CSSTannerGraph(3, [], [[1, 2], [0, 2]])}[code]
    n = t.stgx.nq
    qc, data, b = E.encode_stabilizers(t)
    rng = np.random.default_rng(2)
    p = [list(x / x.sum()) for x in rng.uniform(0.02, 1.0, size=(n, 4)) * np.array([8, 1, 1, 1])]
    measured = sorted(b.ordering[: b.matrix.shape[0]])
    cl = E.clifford_network(qc, n)
    syns = [np.zeros(len(measured), dtype=np.uint8)] + [rng.integers(0, 2, len(measured), dtype=np.uint8) for _ in range(3)]
    for k in range(n):
        f, c, nv, nc, no = E.inference_graph(cl, p, measured, k)
        merged = S.merge_overlapping(f, nv, {v for ch in c for v in ch.vars}, allow_negative=True)
        order = S.choose_order(merged, c)
        for s in syns:
            full = np.concatenate([s, [0]]).astype(np.uint8)
            mar = frontier.run(merged, c, order, 1, full[None, :], nv)[0]
            ref = bf.marginals(qc, n, p, {q: int(s[i]) for i, q in enumerate(measured)})[k]
            got = mar if k in measured else mar[[0, 1, 3, 2]]
            got = got / got.sum()
            assert np.allclose(got, ref, rtol=1e-10, atol=1e-14), (k, s, got, ref)
