import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tensorqec.jl_b200 as tq
t = tq.CSSTannerGraph(tq.SurfaceCode(3, 3))
for name, prob in (("classical", t.stgz), ("css", t)):
    try:
        ct = tq.compile(tq.TNMAP(), prob)
        plan = ct.plan if hasattr(ct, "plan") else ct.cd.plan
        print(name, plan.geometry(), plan.sch.w_max, len(plan.sch.steps))
        if name == "classical":
            syn = tq.SimpleSyndrome(np.zeros((5, 4), dtype=np.uint8))
        else:
            syn = tq.CSSSyndrome(np.zeros((5, 4), dtype=np.uint8), np.zeros((5, 4), dtype=np.uint8))
        print(tq.decode(ct, syn).logp)
    except Exception as e:
        print(name, "ERR", e)
