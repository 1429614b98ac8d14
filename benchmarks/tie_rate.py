#!/usr/bin/env python
"""How often do two exact MAP contractions disagree under the reference's default (p, p, p) noise?  d = 9, p = 0.05:
the dense (reference-style) pairwise contraction vs the frontier recurrence, both in C on the host (oracle/), same
syndromes (Philox seed 9).  Values agree to rounding; the corrections are both maximisers; on exactly tied shots the two
tie-breaks pick different maximisers and, in a fraction of those, different logical classes.  DESIGN.md section 2 quotes
the output of this script."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["TQEC_NO_SWEEP"] = "1"
import tensorqec.jl_b200 as tq  # noqa: E402
from oracle import cref, gf2, networks, philox  # noqa: E402

d, B = 9, int(sys.argv[1]) if len(sys.argv) > 1 else 3072
t = tq.CSSTannerGraph(tq.SurfaceCode(d, d))
em = tq.iid_error(0.05, t)
ex, ez = philox.sample_depolarizing(em.px, em.py, em.pz, 9, 0, B)
sx, sz = gf2.css_syndrome(ex, ez, t.stgx.H, t.stgz.H)
syn = np.concatenate([sx, sz], axis=1)
gdp, _ = tq.reduce2general(t, em)
sch = tq.tnmap_schedule(tq.TNMAP(), gdp)
lp_f, cfg_f = cref.FrontierPlan(sch).run(syn)
nq, s2q, pix, pri = networks.general_problem_css(t, em.px, em.py, em.pz)
lp_d, cfg_d = cref.DensePlan(networks.tnmap_network(nq, s2q, pix, pri), len(s2q), nq, True).run(syn)
lx, lz = tq.logical_operator(t)
n = d * d
diff_class = gf2.check_logical_error_css(cfg_f[:, :n], cfg_f[:, n:], cfg_d[:, :n], cfg_d[:, n:], lx, lz)
ler_f = gf2.check_logical_error_css(ex, ez, cfg_f[:, :n], cfg_f[:, n:], lx, lz)
ler_d = gf2.check_logical_error_css(ex, ez, cfg_d[:, :n], cfg_d[:, n:], lx, lz)
print(json.dumps({"shots": B, "max_rel_value_diff": float(np.max(np.abs(lp_f - lp_d) / np.abs(lp_d))),
                  "identical_corrections": int((cfg_f == cfg_d).all(axis=1).sum()),
                  "different_logical_class": int(diff_class.sum()), "different_logical_class_rate": float(diff_class.mean()),
                  "logical_failures_frontier": int(ler_f.sum()), "logical_failures_dense": int(ler_d.sum())}))
