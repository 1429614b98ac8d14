"""Fused Monte-Carlo pipeline (tqec_mc_run) over chunk sizes: d = 9, p = 0.05, 1e7 shots."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import tensorqec.jl_b200 as tq
from tensorqec.jl_b200 import _cabi, threshold

t = tq.CSSTannerGraph(tq.SurfaceCode(9, 9))
em = tq.iid_error(0.05, t)
mc = threshold.MonteCarlo(t, tq.TNMAP(), em)
for chunk in (1 << 20, 1 << 21, 1 << 22, 10_000_000):
    for rep in range(2):
        counts, ms = mc.run(10_000_000, seed=9, shot_offset=0, chunk=chunk)
    print(json.dumps({"chunk": chunk, "ms": ms, "shots_per_s": 1e7 / ms * 1e3, "counts": [int(c) for c in counts]}), flush=True)
