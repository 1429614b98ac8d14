import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tensorqec.jl_b200 as tq
from tensorqec.jl_b200 import _cabi
from benchmarks.configs import time_map
d, B = 5, 1000000
t = tq.CSSTannerGraph(tq.SurfaceCode(d, d)); em = tq.iid_error(0.05, t)
ct = tq.compile(tq.TNMAP(), t, em)
plan = ct.cd.plan
words = _cabi.sample_errors(_cabi.MODEL_DEPOL, [em.px, em.py, em.pz], 3, 0, B, 0)
n = d * d
H = np.zeros((t.stgx.ns + t.stgz.ns, 2 * n), dtype=np.uint8); H[:t.stgx.ns, n:] = t.stgx.H; H[t.stgx.ns:, :n] = t.stgz.H
syn = _cabi.GF2Matrix(H).apply(words)
ms = time_map(plan, syn)
print(json.dumps({"d": d, "sweep": plan.query(_cabi.Q_SWEEP), "M_per_s": B / ms / 1e3}))
