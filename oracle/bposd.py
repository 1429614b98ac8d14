"""The reference's BP + OSD decoder restated in numpy, one shot at a time (oracle; see oracle/__init__.py).
src/decoding/bposd.jl:55-78 (`belief_propagation`) and :80-97 (`osd`)."""
import numpy as np


def belief_propagation(s2q, q2s, nq, mu, syn, max_iter=100):
    mq2s = {(q, s): mu[q] for q in range(nq) for s in q2s[q]}
    ms2q = {(s, q): 0.0 for s in range(len(s2q)) for q in s2q[s]}
    q_vec = np.zeros(nq)
    with np.errstate(divide="ignore", invalid="ignore"):
        for _ in range(max_iter):
            for s, qs in enumerate(s2q):
                pro = 1.0
                for q in qs:
                    pro *= np.tanh(mq2s[(q, s)])
                for q in qs:
                    ms2q[(s, q)] = (-1.0) ** int(syn[s]) * np.arctanh(pro / np.tanh(mq2s[(q, s)]))
            for q in range(nq):
                q_vec[q] = sum(ms2q[(s, q)] for s in q2s[q]) + mu[q]
                for s in q2s[q]:
                    mq2s[(q, s)] = max(min(q_vec[q] - ms2q[(s, q)], 10.0), -10.0)
            e = (q_vec < 0).astype(np.uint8)
            if all((sum(e[q] for q in qs) & 1) == int(syn[s]) for s, qs in enumerate(s2q)):
                return True, e, np.argsort(q_vec, kind="stable")
    return False, np.zeros(nq, dtype=np.uint8), np.argsort(q_vec, kind="stable")


def osd(H, order, syn):
    """bposd.jl:80-97: the first linearly independent columns of H in `order` (the reference starts its list with
    order[1] twice; the duplicate is dependent and drops out), solve for the syndrome on those columns."""
    H = np.asarray(H, dtype=np.uint8)
    ns, nq = H.shape
    rows = H.copy()
    rhs = np.asarray(syn, dtype=np.uint8).copy()
    used = np.zeros(ns, dtype=bool)
    piv = {}
    for q in order:
        if len(piv) == ns:
            break
        cand = [s for s in range(ns) if not used[s] and rows[s, q]]
        if not cand:
            continue
        r = cand[0]
        used[r] = True
        piv[r] = q
        for s in range(ns):
            if s != r and rows[s, q]:
                rows[s] ^= rows[r]
                rhs[s] ^= rhs[r]
    e = np.zeros(nq, dtype=np.uint8)
    for r, q in piv.items():
        e[q] = rhs[r]
    return e


def decode(H, s2q, q2s, p, syn, max_iter=100, use_osd=True):
    nq = H.shape[1]
    mu = np.log((1 - np.asarray(p)) / np.asarray(p))
    ok, e, order = belief_propagation(s2q, q2s, nq, mu, syn, max_iter)
    if ok or not use_osd:
        return ok, e, ok
    return True, osd(H, order, syn), False
