import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tensorqec.jl_b200 as tq
from oracle import cref
t = tq.CSSTannerGraph(tq.SurfaceCode(3, 3))
ct = tq.compile(tq.TNMAP(), t)
plan = ct.cd.plan
sch = plan.sch
print(plan.geometry(), [(s.w_in, s.w_out, len(s.opened), len(s.closed), len(s.ker)) for s in sch.steps])
syn = ((np.arange(64)[:, None] >> np.arange(8)) & 1).astype(np.uint8)
corr, lp = plan.decode_map(tq.pack_bits(syn))
lp2, cfg2 = cref.FrontierPlan(sch).run(syn)
print("logp equal:", np.array_equal(lp, lp2), " cfg equal:", np.array_equal(tq.unpack_bits(corr, 18), cfg2))
bad = np.flatnonzero(lp != lp2)
print("bad shots", bad[:20], lp[bad[:6]], lp2[bad[:6]])
