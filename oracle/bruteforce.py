"""Exhaustive enumeration for small codes (oracle tier 1; see oracle/__init__.py).

Semantic definition of what the decoders compute: over all 2^nq assignments e of the error
variables, weight(e) = prod_f prior_f(e|vars_f); TNMAP = argmax over {e : H e = s}; TNMMAP = sum of weight over
{e : H e = s, L e = l} for every logical sector l.  Feasible for nq <= ~20 (d=3 surface: 2^18; Steane: 2^14).
"""
import numpy as np


def all_assignments(nq):
    a = np.arange(1 << nq, dtype=np.int64)
    return ((a[:, None] >> np.arange(nq)) & 1).astype(np.uint8)          # (2^nq, nq), bit v of row a = var v


def weights(nq, prior_ixs, priors, E=None):
    """-> (E, w, logw): weight of every assignment, product / log-sum taken in the order the priors are listed."""
    E = all_assignments(nq) if E is None else E
    w = np.ones(E.shape[0])
    logw = np.zeros(E.shape[0])
    with np.errstate(divide="ignore"):
        for ix, t in zip(prior_ixs, priors):
            t = np.asarray(t, dtype=np.float64)
            v = t[tuple(E[:, l] for l in ix)]
            w = w * v
            logw = logw + np.log(v)
    return E, w, logw


def syndromes_of(E, s2q):
    S = np.zeros((E.shape[0], len(s2q)), dtype=np.uint8)
    for s, c in enumerate(s2q):
        for q in c:
            S[:, s] ^= E[:, q]
    return S


class Enumeration:
    """All assignments of a general decoding problem, grouped by syndrome."""

    def __init__(self, nq, s2q, prior_ixs, priors):
        self.nq = nq
        self.E, self.w, self.logw = weights(nq, prior_ixs, priors)
        S = syndromes_of(self.E, s2q)
        self.key = (S.astype(np.int64) << np.arange(len(s2q))).sum(axis=1)
        self.ns = len(s2q)

    def _sel(self, syndrome):
        k = int((np.asarray(syndrome, dtype=np.int64) << np.arange(self.ns)).sum())
        return np.flatnonzero(self.key == k)

    def map(self, syndrome, rtol=1e-12):
        """-> (max log-weight, maximisers (m, nq)); ties decided on the exact product within rtol."""
        idx = self._sel(syndrome)
        if idx.size == 0 or self.w[idx].max() == 0.0:
            return -np.inf, np.zeros((0, self.nq), dtype=np.uint8)
        wm = self.w[idx].max()
        best = idx[self.w[idx] >= wm * (1 - rtol)]
        return float(self.logw[best].max()), self.E[best]

    def marginal(self, syndrome, L):
        """-> array over the 2^len(L) sectors, sector index = sum_i parity(L[i].e) << i."""
        idx = self._sel(syndrome)
        L = np.asarray(L, dtype=np.int64)
        par = (self.E[idx].astype(np.int64) @ L.T) & 1
        sec = (par << np.arange(L.shape[0])).sum(axis=1)
        out = np.zeros(1 << L.shape[0])
        np.add.at(out, sec, self.w[idx])
        return out
