"""Monte-Carlo driver: the caller whose inner body is the whole hot path.

Reference: src/decoding/threshold.jl:1-19 (`multi_round_qec`: compile once; per round sample -> syndrome -> decode ->
logical check -> three counters) and :21-33 (classical variant).  Here the loop body is ONE fused device pipeline
(`tqec_mc_run`): Philox sampling, popcount-parity syndrome extraction, the frontier decoder and the logical check
run back to back on the GPU over chunks of shots; only four int64 counters come back.

Logical check: the reference's loop passes stabilizer matrices where logical operators belong and inverts the sense
(SURVEY D.1); the counters here use the sound recipe the reference documents elsewhere
(`logical_operator` + `check_logical_error(e, e_dec, lx, lz)`, docs/src/tndecoder.md:72-78).
"""
from __future__ import annotations

import numpy as np

from . import _cabi
from .decoding import TNMAP, compile
from .error_model import IndependentDepolarizingError, IndependentFlipError, iid_error
from .tanner import CSSTannerGraph, SimpleTannerGraph, logical_operator


def css_general_matrices(tanner: CSSTannerGraph, lx, lz):
    """Check and logical matrices over the 2n general variables (x block, then z block; reduce2general order):
    H = [X checks on the z block ; Z checks on the x block], L = [lz rows on the x block (class 0: X-type logical
    flips) ; lx rows on the z block (class 1: Z-type)]."""
    n = tanner.stgx.nq
    H = np.zeros((tanner.stgx.ns + tanner.stgz.ns, 2 * n), dtype=np.uint8)
    H[: tanner.stgx.ns, n:] = tanner.stgx.H
    H[tanner.stgx.ns:, :n] = tanner.stgz.H
    L = np.zeros((lz.shape[0] + lx.shape[0], 2 * n), dtype=np.uint8)
    L[: lz.shape[0], :n] = lz
    L[lz.shape[0]:, n:] = lx
    return H, L, [0] * lz.shape[0] + [1] * lx.shape[0]


class MonteCarlo:
    """Compiled sample -> syndrome -> decode -> check pipeline for one (code, noise, decoder, device)."""

    def __init__(self, tanner, decoder, em=None, logical=None, device=None):
        if not isinstance(decoder, TNMAP):
            raise TypeError("the fused Monte-Carlo pipeline drives the TNMAP decoder")
        device = decoder.device if device is None else device
        decoder = TNMAP(decoder.optimizer, device, decoder.table_bits, decoder.head_bits)
        self.device = device
        if isinstance(tanner, CSSTannerGraph):
            em = iid_error(0.05, tanner) if em is None else em
            if not isinstance(em, IndependentDepolarizingError):
                raise TypeError("a CSS code needs an IndependentDepolarizingError")
            ct = compile(decoder, tanner, em)
            self.compiled = ct
            self.plan = ct.cd.plan
            lx, lz = logical_operator(tanner) if logical is None else logical
            H, L, self.row_class = css_general_matrices(tanner, lx, lz)
            self.model, self.probs = _cabi.MODEL_DEPOL, [em.px, em.py, em.pz]
        elif isinstance(tanner, SimpleTannerGraph):
            em = iid_error(0.05, tanner) if em is None else em
            if not isinstance(em, IndependentFlipError):
                raise TypeError("a classical code needs an IndependentFlipError")
            if logical is None:
                raise ValueError("a classical Monte-Carlo run needs the logical check matrix")
            ct = compile(decoder, tanner, em)
            self.compiled = ct
            self.plan = ct.plan
            H = tanner.H
            L = np.asarray(logical, dtype=np.uint8)
            self.row_class = [0] * L.shape[0]
            self.model, self.probs = _cabi.MODEL_FLIP, [em.p]
        else:
            raise TypeError("tanner must be a SimpleTannerGraph or a CSSTannerGraph")
        self.H = _cabi.GF2Matrix(H, device)
        self.L = _cabi.GF2Matrix(L, device)

    def run(self, shots: int, seed: int = 0, shot_offset: int = 0, chunk: int = 0, comm=None):
        """-> (counts[4] = {X-type failures, Z-type failures, any failure, shots}, device milliseconds).  With `comm`
        (a `_cabi.Comm`) the counters are summed over the job's ranks inside the pipeline (one ncclAllReduce)."""
        return _cabi.mc_run(self.plan, self.H, self.L, self.row_class, self.model, self.probs, seed, shot_offset,
                            int(shots), chunk, comm)


def _unfused_css(tanner, decoder, em, rounds, seed, reference_prior):
    """Any other decoder the package compiles for a CSS code (TNMMAP, TableDecoder, ...): sample -> syndrome -> decode ->
    check as separate batched calls (each on the GPU), in chunks."""
    from .error_model import check_logical_error, random_error_pattern, syndrome_extraction
    from .decoding import compile as _compile, decode as _decode
    ct = _compile(decoder, tanner) if reference_prior else _compile(decoder, tanner, em)
    lx, lz = logical_operator(tanner)
    counts = np.zeros(4, dtype=np.int64)
    chunk = 1 << 18
    for lo in range(0, rounds, chunk):
        nb = min(chunk, rounds - lo)
        ep = random_error_pattern(em, seed=seed, shots=nb, shot_offset=lo)
        res = _decode(ct, syndrome_extraction(ep, tanner))
        fx = check_logical_error(ep.xerror, res.error_pattern.xerror, lz)
        fz = check_logical_error(ep.zerror, res.error_pattern.zerror, lx)
        counts += [int(fx.sum()), int(fz.sum()), int((fx | fz).sum()), nb]
    return counts


def multi_round_qec(tanner, decoder, em, tanner_check=None, *, rounds: int = 10, seed: int = 0, device=None,
                    reference_prior: bool = False):
    """threshold.jl:1-19 -> (logical_xerror/rounds, logical_zerror/rounds, logical_error/rounds) for a CSS code;
    threshold.jl:21-33 -> logical_xerror/rounds for a classical code checked against `tanner_check.H`.
    TNMAP runs the fused device pipeline (`tqec_mc_run`); any other decoder runs the same four stages as separate
    batched calls.  The decoder is compiled for the sampling model `em`; `reference_prior=True` reproduces the
    reference, which compiles with `iid_error(0.05)` whatever `em` is (threshold.jl:5: `compile(decoder, tanner)`)."""
    if isinstance(tanner, CSSTannerGraph) and not isinstance(decoder, TNMAP):
        counts = _unfused_css(tanner, decoder, em, rounds, seed, reference_prior)
        return counts[0] / rounds, counts[1] / rounds, counts[2] / rounds
    if reference_prior and isinstance(tanner, CSSTannerGraph):
        mc = MonteCarlo(tanner, decoder, iid_error(0.05, tanner), device=device)
        mc.model, mc.probs = _cabi.MODEL_DEPOL, [em.px, em.py, em.pz]     # sample from `em`, decode with the default prior
        counts, _ = mc.run(rounds, seed)
        return counts[0] / rounds, counts[1] / rounds, counts[2] / rounds
    if isinstance(tanner, CSSTannerGraph):
        counts, _ = MonteCarlo(tanner, decoder, em, device=device).run(rounds, seed)
        return counts[0] / rounds, counts[1] / rounds, counts[2] / rounds
    if tanner_check is None:
        raise TypeError("multi_round_qec(tanner::SimpleTannerGraph, decoder, em, tanner_check)")
    counts, _ = MonteCarlo(tanner, decoder, em, logical=tanner_check.H, device=device).run(rounds, seed)
    return counts[2] / rounds
