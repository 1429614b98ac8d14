"""Pairwise contraction of the reference's dense networks (oracle; see oracle/__init__.py).

Restates what `ct.code(ct.tensors...)` (OMEinsum `NestedEinsum`, tndecoder.jl:162, 250) and
`most_probable_config` (TensorInference, tndecoder.jl:55) compute, per SURVEY B.2/B.4 [3P-recollection]:
  * a binary contraction tree fixed at compile time; here a deterministic greedy optimiser (OMEinsum's
    `GreedyMethod` flavour: repeatedly contract the pair minimising size(out) - size(a) - size(b));
  * sum-product: plain Float64 einsum per step;
  * max-plus (MAP): tensors -> log, forward pass caching every intermediate, then a root-to-leaves traceback that
    picks, per node, the first (column-major, first contracted label fastest) maximiser of lhs + rhs for the
    already-fixed outer labels, both children resolved from the same choice (tie rule B.3).
Evidence labels keep their axis with size 1 (the tensor is sliced), as TensorInference does.
"""
from dataclasses import dataclass
from typing import List, Optional

import numpy as np


@dataclass
class Tree:
    """Binary contraction tree in execution order: step k contracts nodes a, b into node (n_leaves + k)."""
    n_leaves: int
    steps: List[tuple]                 # (a, b, out_labels, contracted_labels)
    labels: List[List[int]]            # labels of every node (leaves first)

    def complexity(self, sizes=None):
        tc = 0.0
        sc = 0
        rw = 0.0
        for a, b, out, con in self.steps:
            la, lb = self.labels[a], self.labels[b]
            un = set(la) | set(lb)
            dim = lambda ls: float(np.prod([sizes.get(l, 2) if sizes else 2 for l in ls])) if ls else 1.0
            tc += dim(un)
            rw += dim(la) + dim(lb) + dim(out)
            sc = max(sc, int(round(np.log2(max(dim(out), 1.0)))))
        return {"log2_tc": float(np.log2(tc)) if tc else 0.0, "sc": sc, "log2_rw": float(np.log2(rw)) if rw else 0.0}


def greedy_tree(ixs, iy, sizes=None) -> Tree:
    """Deterministic greedy pairwise order.  `sizes[label]` (default 2) lets evidence labels count as size 1."""
    sizes = sizes or {}
    lg = lambda ls: sum(0 if sizes.get(l, 2) == 1 else 1 for l in ls)
    labels = [list(ix) for ix in ixs]
    n = len(ixs)
    alive = set(range(n))
    count = {}
    for ix in ixs:
        for l in set(ix):
            count[l] = count.get(l, 0) + 1
    for l in iy:
        count[l] = count.get(l, 0) + 1                     # open labels are never summed
    holders = {}
    for t, ix in enumerate(ixs):
        for l in ix:
            holders.setdefault(l, set()).add(t)
    steps = []

    def out_of(a, b):
        la, lb = labels[a], labels[b]
        un = list(dict.fromkeys(la + lb))
        out, con = [], []
        for l in un:
            k = (l in la) + (l in lb)
            (con if count[l] == k else out).append(l)
        return out, con

    while len(alive) > 1:
        best = None
        seen = set()
        for l, hs in holders.items():
            hl = sorted(hs)
            for i in range(len(hl)):
                for j in range(i + 1, len(hl)):
                    p = (hl[i], hl[j])
                    if p in seen:
                        continue
                    seen.add(p)
                    out, con = out_of(*p)
                    loss = 2.0 ** lg(out) - 2.0 ** lg(labels[p[0]]) - 2.0 ** lg(labels[p[1]])
                    key = (loss, lg(out), p)
                    if best is None or key < best[0]:
                        best = (key, p, out, con)
        if best is None:                                    # disconnected pieces: outer product of the two smallest
            hl = sorted(alive, key=lambda t: (lg(labels[t]), t))[:2]
            p = (min(hl), max(hl))
            out, con = out_of(*p)
            best = (None, p, out, con)
        _, (a, b), out, con = best
        new = len(labels)
        labels.append(out)
        steps.append((a, b, out, con))
        for l in set(labels[a]) | set(labels[b]):
            k = (l in labels[a]) + (l in labels[b])
            holders[l].discard(a)
            holders[l].discard(b)
            if l in out:
                holders[l].add(new)
                count[l] = count[l] - k + 1
            else:
                del holders[l]
        alive -= {a, b}
        alive.add(new)
    return Tree(n, steps, labels)


# ------------------------------------------------------------------------------------------------------------
def _slice_evidence(net):
    """Slice every tensor on its evidence labels with a length-1 range (label kept, size 1)."""
    out = []
    for ix, t in zip(net.ixs, net.tensors):
        t = np.asarray(t, dtype=np.float64)
        for ax, l in enumerate(ix):
            if l in net.evidence:
                t = np.take(t, [net.evidence[l]], axis=ax)
        out.append(t)
    return out


def _pair(ta, la, tb, lb, out, con, maxplus):
    """One pairwise step on named axes: out[out] = (+)_con ta[la] (x) tb[lb]."""
    un = out + con
    pos = {l: i for i, l in enumerate(un)}

    def expand(t, ls):
        perm = sorted(range(len(ls)), key=lambda i: pos[ls[i]])
        t = np.transpose(t, perm)
        shape = [1] * len(un)
        for i in perm:
            shape[pos[ls[i]]] = t.shape[perm.index(i)]
        return t.reshape(shape)

    A, B = expand(ta, la), expand(tb, lb)
    red = tuple(range(len(out), len(un)))
    if maxplus:
        full = A + B
        return full.max(axis=red) if red else full
    full = A * B
    return full.sum(axis=red) if red else full


def contract_sumproduct(net, tree: Optional[Tree] = None):
    """-> ndarray with one axis per label of net.iy (in that order)."""
    tensors = _slice_evidence(net)
    sizes = {l: 1 for l in net.evidence}
    tree = tree or greedy_tree(net.ixs, net.iy, sizes)
    vals = list(tensors)
    for a, b, out, con in tree.steps:
        vals.append(_pair(vals[a], tree.labels[a], vals[b], tree.labels[b], out, con, False))
    res, lab = vals[-1], tree.labels[-1]
    if len(net.ixs) == 1:
        res, lab = vals[0], list(net.ixs[0])
    perm = [lab.index(l) for l in net.iy]
    return np.transpose(res, perm)


def most_probable_config(net, tree: Optional[Tree] = None):
    """-> (logp, config) with config[v] for every label v in 0..nvars-1 (evidence labels report their value)."""
    tensors = _slice_evidence(net)
    sizes = {l: 1 for l in net.evidence}
    tree = tree or greedy_tree(net.ixs, net.iy, sizes)
    with np.errstate(divide="ignore"):
        vals = [np.log(t) for t in tensors]
    for a, b, out, con in tree.steps:
        vals.append(_pair(vals[a], tree.labels[a], vals[b], tree.labels[b], out, con, True))
    root = len(vals) - 1
    logp = float(vals[root].reshape(-1)[0]) if vals[root].size == 1 else float(vals[root].max())
    assign = dict(net.evidence)                          # label -> value (index into the *unsliced* axis)

    def axis_index(l):
        return 0 if l in net.evidence else assign[l]

    # walk the steps backwards: when a node is visited all of its own labels are assigned
    for k in range(len(tree.steps) - 1, -1, -1):
        a, b, out, con = tree.steps[k]
        if not con:
            continue
        la, lb = tree.labels[a], tree.labels[b]

        def sub(t, ls):
            idx = tuple(slice(None) if l in con else axis_index(l) for l in ls)
            kept = [l for l in ls if l in con]
            return t[idx], kept

        sa, ka = sub(vals[a], la)
        sb, kb = sub(vals[b], lb)
        s = _pair(sa, ka, sb, kb, list(con), [], True)   # axes ordered as `con`
        flat = s.reshape(-1, order="F")                  # first contracted label fastest
        j = int(np.argmax(flat))                         # first maximiser
        for i, l in enumerate(con):
            dim = s.shape[i]
            assign[l] = 0 if l in net.evidence else j % dim
            j //= dim
    config = np.array([assign.get(v, 0) for v in range(net.nvars)], dtype=np.uint8)
    return logp, config
