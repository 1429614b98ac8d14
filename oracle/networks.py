"""The reference's tensor networks, restated label for label (oracle; see oracle/__init__.py).

A network is (ixs, tensors, iy, evidence): `ixs[t]` = labels of tensor t (0-based ints), `tensors[t]` = ndarray with
one axis per label IN THAT ORDER (so `tensors[t][i0, i1, ...]` is the reference's `T[i0+1, i1+1, ...]`), `iy` = open
labels of the result, `evidence` = {label: clamped value} (TensorInference-style: the label stays, with size 1).

Reference:
  parity_check_matrix            src/decoding/tndecoder.jl:42-44    1.0 iff the (k+1)-bit index has even popcount
  single_qubit_tensor            src/decoding/general_decoding.jl:5 [1-px-py-pz  pz ; px  py]  indexed [x, z]
  reduce2general                 src/decoding/general_decoding.jl:24-32
  stg2uaimodel + TNMAP compile   src/decoding/tndecoder.jl:33-50    (+ per-variable ones(2), SURVEY B.1 [3P])
  TNMMAP CSS compile             src/decoding/tndecoder.jl:97-146
  TNMMAP DEM compile             src/decoding/tndecoder.jl:186-238  (`push_check_node!` chain factorisation)
"""
from dataclasses import dataclass, field
from typing import Dict, List

import numpy as np


def parity_check_matrix(k: int) -> np.ndarray:
    """tndecoder.jl:42-44: rank k+1, entry 1.0 iff popcount(linear index) is even."""
    idx = np.arange(1 << (k + 1))
    pop = np.zeros_like(idx)
    for b in range(k + 1):
        pop += (idx >> b) & 1
    return (1.0 - (pop % 2)).reshape((2,) * (k + 1), order="F")


def single_qubit_tensor(px, py, pz) -> np.ndarray:
    """general_decoding.jl:5: T[x, z]."""
    return np.array([[1.0 - px - py - pz, pz], [px, py]])


@dataclass
class Network:
    ixs: List[List[int]]
    tensors: List[np.ndarray]
    iy: List[int]
    evidence: Dict[int, int] = field(default_factory=dict)
    nvars: int = 0
    ev_bit: Dict[int, int] = field(default_factory=dict)      # evidence label -> syndrome bit that clamps it
    syn_leaf: Dict[int, int] = field(default_factory=dict)    # leaf index of a rank-1 syndrome vector -> syndrome bit


def general_problem_css(tanner, px, py, pz):
    """reduce2general (general_decoding.jl:24-32): 2n variables (x_i = i, z_i = i + n); checks = X checks on the
    z block, then Z checks on the x block; one 2x2 prior per qubit on (i, i+n).  -> (nq, s2q, prior_ixs, priors)"""
    n = tanner.stgx.nq
    s2q = [[q + n for q in s] for s in tanner.stgx.s2q] + [list(s) for s in tanner.stgz.s2q]
    ixs = [[i, i + n] for i in range(n)]
    pri = [single_qubit_tensor(px[i], py[i], pz[i]) for i in range(n)]
    return 2 * n, s2q, ixs, pri


def general_problem_classical(tanner, p):
    """interfaces.jl:139-142: rank-1 priors [1-p, p] on each bit."""
    return tanner.nq, [list(s) for s in tanner.s2q], [[i] for i in range(tanner.nq)], [np.array([1 - q, q]) for q in p]


def tnmap_network(nq, s2q, prior_ixs, priors, syndrome=None) -> Network:
    """stg2uaimodel + TensorNetworkModel (tndecoder.jl:33-50): variables 0..nq-1, syndrome variable of check s is
    nq+s; tensors = ones(2) per variable, then the parity factors on (s2q[s]..., nq+s), then the priors;
    evidence clamps every syndrome variable; no open label."""
    ns = len(s2q)
    nvars = nq + ns
    ixs = [[v] for v in range(nvars)]
    tensors = [np.ones(2) for _ in range(nvars)]
    for s, c in enumerate(s2q):
        ixs.append(list(c) + [nq + s])
        tensors.append(parity_check_matrix(len(c)))
    for ix, t in zip(prior_ixs, priors):
        ixs.append(list(ix))
        tensors.append(np.asarray(t, dtype=np.float64))
    ev = {nq + s: int(0 if syndrome is None else syndrome[s]) for s in range(ns)}
    return Network(ixs, tensors, [], ev, nvars, {nq + s: s for s in range(ns)})


def tnmmap_css_network(tanner, lx, lz, px, py, pz, sx=None, sz=None) -> Network:
    """tndecoder.jl:97-146.  Labels: x errors 0..n-1, z errors n..2n-1, X-check syndromes, Z-check syndromes,
    lx-parities (of Z errors), lz-parities (of X errors).  Output = the 2k logical labels (lx block first)."""
    n = tanner.stgx.nq
    nsx, nsz = tanner.stgx.ns, tanner.stgz.ns
    nsyn = nsx + nsz
    k = lx.shape[0]
    nvars = 2 * n + nsyn + 2 * k
    ixs, tensors = [], []
    for i, c in enumerate(tanner.stgx.s2q):
        ixs.append([q + n for q in c] + [2 * n + i])
        tensors.append(parity_check_matrix(len(c)))
    for i, c in enumerate(tanner.stgz.s2q):
        ixs.append(list(c) + [2 * n + nsx + i])
        tensors.append(parity_check_matrix(len(c)))
    for i in range(n):
        ixs.append([i, i + n])
        tensors.append(single_qubit_tensor(px[i], py[i], pz[i]))
    for i in range(k):
        sup = [int(q) for q in np.flatnonzero(lz[i])]
        ixs.append(sup + [2 * n + nsyn + k + i])
        tensors.append(parity_check_matrix(len(sup)))
    for i in range(k):
        sup = [int(q) + n for q in np.flatnonzero(lx[i])]
        ixs.append(sup + [2 * n + nsyn + i])
        tensors.append(parity_check_matrix(len(sup)))
    syn = np.zeros(nsyn, dtype=int)
    if sx is not None:
        syn[:nsx] = sx
    if sz is not None:
        syn[nsx:] = sz
    syn_leaf = {}
    for j in range(nsyn):                       # rank-1 syndrome vectors (update_syndrome!, :148-158)
        syn_leaf[len(ixs)] = j
        ixs.append([2 * n + j])
        tensors.append(np.array([0.0, 1.0]) if syn[j] else np.array([1.0, 0.0]))
    iy = list(range(2 * n + nsyn, nvars))
    return Network(ixs, tensors, iy, {}, nvars, {}, syn_leaf)


def _push_check_node(ixs, tensors, c, check_label, nvars, factorize):
    """tndecoder.jl:221-238 (the weight-1 chain case indexes c[2] and throws in the reference, SURVEY D.4;
    here it falls back to the unfactorised tensor)."""
    L = len(c)
    if (not factorize) or L <= 2:
        ixs.append(list(c) + [check_label])
        tensors.append(parity_check_matrix(L))
        return nvars
    ixs.append([c[0], c[1], nvars])
    tensors.append(parity_check_matrix(2))
    for j in range(2, L - 1):                   # Julia j = 3 .. L-1
        ixs.append([nvars + j - 1, c[j], nvars + j - 2])
        tensors.append(parity_check_matrix(2))
    ixs.append([check_label, c[L - 1], nvars + L - 3])
    tensors.append(parity_check_matrix(2))
    return nvars + L - 2


def tnmmap_dem_network(error_rates, flipped, n_det, n_obs, syndrome=None, factorize=True) -> Network:
    """tndecoder.jl:186-219.  `flipped[e]` = 0-based detector ids (< n_det) and observable ids (n_det + l).
    Labels: mechanisms, detectors, observables, then auxiliaries."""
    ne = len(error_rates)
    nvars = ne + n_det + n_obs
    iy = list(range(ne + n_det, nvars))
    ixs, tensors = [], []
    for d in range(n_det):
        c = [e for e in range(ne) if d in flipped[e]]
        nvars = _push_check_node(ixs, tensors, c, ne + d, nvars, factorize)
    for l in range(n_obs):
        c = [e for e in range(ne) if (n_det + l) in flipped[e]]
        nvars = _push_check_node(ixs, tensors, c, ne + n_det + l, nvars, factorize)
    for e, p in enumerate(error_rates):
        ixs.append([e])
        tensors.append(np.array([1.0 - p, p]))
    syn_leaf = {}
    for d in range(n_det):
        syn_leaf[len(ixs)] = d
        ixs.append([ne + d])
        s = 0 if syndrome is None else int(syndrome[d])
        tensors.append(np.array([0.0, 1.0]) if s else np.array([1.0, 0.0]))
    return Network(ixs, tensors, iy, {}, nvars, {}, syn_leaf)
