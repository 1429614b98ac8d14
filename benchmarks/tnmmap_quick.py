"""TNMMAP (CSS) throughput, d = 5 / 7 / 9, sweep vs general kernels (TQEC_NO_SWEEP=1): one line per case."""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tensorqec.jl_b200 as tq
from tensorqec.jl_b200 import _cabi
from benchmarks.configs import time_marginal

CASES = ((int(sys.argv[1]), int(sys.argv[2])),) if len(sys.argv) > 2 else ((5, 1000000), (7, 1000000), (9, 400000))
for d, B in CASES:
    t = tq.CSSTannerGraph(tq.SurfaceCode(d, d))
    em = tq.iid_error(0.05, t)
    ct = tq.compile(tq.TNMMAP(), t, em)
    n = d * d
    words = _cabi.sample_errors(_cabi.MODEL_DEPOL, [em.px, em.py, em.pz], 3, 0, B, 0)
    H = np.zeros((t.stgx.ns + t.stgz.ns, 2 * n), dtype=np.uint8)
    H[:t.stgx.ns, n:] = t.stgx.H
    H[t.stgx.ns:, :n] = t.stgz.H
    syn = _cabi.GF2Matrix(H).apply(words)
    ms = time_marginal(ct.plan, syn)
    print(json.dumps({"case": f"TNMMAP d={d} CSS", "shots": B, "ms": ms, "syndromes_per_s": B / ms * 1e3,
                      "sweep": ct.plan.query(_cabi.Q_SWEEP), "w_max": ct.schedule.w_max}))
