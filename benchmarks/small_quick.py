"""Tabulated plans (n_checks <= 16): throughput and single-shot latency, table vs kernels (TQEC_NO_TABLE=1)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tensorqec.jl_b200 as tq
from tensorqec.jl_b200 import _cabi
from benchmarks.configs import time_map
for name, code in (("surface3", tq.SurfaceCode(3, 3)), ("steane", tq.SteaneCode()), ("color488_5", tq.Color488(5))):
    t = tq.CSSTannerGraph(code); em = tq.iid_error(0.05, t)
    ct = tq.compile(tq.TNMAP(), t, em)
    plan = ct.cd.plan
    n = t.stgx.nq
    H = np.zeros((t.stgx.ns + t.stgz.ns, 2 * n), dtype=np.uint8); H[:t.stgx.ns, n:] = t.stgx.H; H[t.stgx.ns:, :n] = t.stgz.H
    for B in (1, 1000000, 10000000):
        words = _cabi.sample_errors(_cabi.MODEL_DEPOL, [em.px, em.py, em.pz], 3, 0, B, 0)
        syn = _cabi.GF2Matrix(H).apply(words)
        ms = time_map(plan, syn)
        print(json.dumps({"code": name, "B": B, "table": plan.query(_cabi.Q_TABLE), "us": ms * 1e3, "M_per_s": B / ms / 1e3}))
