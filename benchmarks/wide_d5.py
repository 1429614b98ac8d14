#!/usr/bin/env python
"""BASELINE configs[3], d = 5 x 5 rounds: circuit-level rotated-surface memory DEM -> TNMMAP through the global-memory
executor.  Prints plan statistics, decode time per shot, achieved HBM bandwidth and the marginals (JSON lines)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tensorqec.jl_b200 as tq  # noqa: E402
from tensorqec.jl_b200 import _cabi  # noqa: E402


def main():
    d = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    p = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-3
    txt = tq.surface_memory_circuit(d, d, "Z", p, p, p, p)
    dem = tq.detector_error_model(tq.parse_stim_string(txt))
    t0 = time.time()
    ct = tq.compile(tq.TNMMAP(), dem)
    t_compile = time.time() - t0
    wp = ct.schedule
    info = dict(d=d, mechanisms=len(dem.error_rates), detectors=dem.n_detectors, compile_s=round(t_compile, 2),
                wide=ct.plan.query(_cabi.Q_WIDE))
    if hasattr(wp, "passes"):
        info.update(passes=len(wp.passes), w_cap=wp.w_cap, candidates=wp.cost, bytes_per_shot=wp.bytes_per_shot)
    print(json.dumps(info), flush=True)
    ep = tq.random_error_pattern(dem, seed=55, shots=B)
    syn = tq.syndrome_extraction(ep, ct.tanner)
    for rep in range(2):
        t0 = time.time()
        res = tq.decode(ct, syn)
        dt = time.time() - t0
        line = dict(rep=rep, shots=B, s_per_shot=dt / B, batch=ct.plan.query(_cabi.Q_WIDE_BATCH))
        if hasattr(wp, "passes"):
            line.update(hbm_gbs=wp.bytes_per_shot * B / dt / 1e9, gcand_per_s=wp.cost * B / dt / 1e9)
        print(json.dumps(line), flush=True)
    true_obs = (ep[:, ct.l2q[0]].sum(axis=1) & 1)
    print(json.dumps(dict(marginal=res.marginal.reshape(B, -1, order="F").tolist(), sector=res.sector.tolist(),
                          true_obs=true_obs.tolist(), syndrome_weight=syn.s.sum(axis=1).tolist())), flush=True)


if __name__ == "__main__":
    main()
