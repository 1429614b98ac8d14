#!/usr/bin/env python
"""DEM TNMMAP plans that fit on chip, decoded by the on-chip general kernels and by the global-memory executor's tile
kernel (TQEC_FORCE_WIDE): which one should such plans use?  One JSON line per (case, path)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tensorqec.jl_b200 as tq            # noqa: E402
from tensorqec.jl_b200 import _cabi      # noqa: E402


def time_marginal(plan, words, reps=3):
    B = words.shape[0]
    d_syn = torch.from_numpy(words.view(np.int64)).cuda()
    d_mar = torch.empty((B, 1 << plan.n_obs), dtype=torch.float64, device="cuda")
    d_arg = torch.empty((B,), dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream()
    for _ in range(2):
        plan.decode_marginal_dev(d_syn.data_ptr(), B, d_mar.data_ptr(), d_arg.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        plan.decode_marginal_dev(d_syn.data_ptr(), B, d_mar.data_ptr(), d_arg.data_ptr(), st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, d_mar.cpu().numpy()


def main():
    cases = [("circuit d=3x3", tq.detector_error_model(tq.parse_stim_string(tq.surface_memory_circuit(3, 3, "Z", 1e-3, 1e-3, 1e-3, 1e-3))), 100_000),
             ("phenom d=5x5", tq.parse_dem_file(os.path.join(ROOT, "tests", "golden", "surface_d5_r5_phenom.dem")), 100_000),
             ("phenom d=3x3", tq.parse_dem_file(os.path.join(ROOT, "tests", "golden", "surface_d3_r3_phenom.dem")), 1_000_000),
             ("colour d=3 r=2", tq.parse_dem_file(os.path.join(ROOT, "tests", "golden", "color_memory_xyz_d3_r2.dem")), 1_000_000)]
    for name, dem, B in cases:
        ref = None
        for path in ("on-chip", "wide"):
            os.environ["TQEC_SUMPROD_ONCHIP_WIDTH"] = "13"        # "on-chip": never the global-memory executor below 14 bits
            if path == "wide":
                os.environ["TQEC_FORCE_WIDE"] = "1"
            ct = tq.compile(tq.TNMMAP(table_bits=0), dem)
            os.environ.pop("TQEC_FORCE_WIDE", None)
            os.environ.pop("TQEC_SUMPROD_ONCHIP_WIDTH", None)
            ep = _cabi.sample_errors(_cabi.MODEL_FLIP, [np.asarray(dem.error_rates)], 3, 0, B)
            syn = _cabi.GF2Matrix(ct.tanner.H).apply(ep)
            ms, mar = time_marginal(ct.plan, syn)
            same = None if ref is None else bool(np.allclose(mar, ref, rtol=1e-10, atol=0))
            ref = mar if ref is None else ref
            print(json.dumps({"case": name, "path": path, "shots": B, "ms": ms, "syndromes_per_s": B / (ms * 1e-3),
                              "lowered": ct.plan.lowered, "matches_on_chip": same}), flush=True)


if __name__ == "__main__":
    main()
