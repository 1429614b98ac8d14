#!/usr/bin/env python
"""All BASELINE.json configs other than the headline one (which bench.py measures): one JSON line per case.

  1. d=3 surface, TNMAP, p=0.05, 1000 syndromes                       (configs[0])
  2. d=5 / d=7 surface, TNMAP, p = 0.01..0.10 sweep, 1e6 syndromes     (configs[1])
  4. DEM TNMMAP on the reference's DEM fixture                          (configs[3]; stim-generated surface-memory DEMs
                                                                         need stim or the circuit->DEM generator)
  5. Color488(5) and Steane, TNMAP, batch-size sweep 1..1e6            (configs[4]: latency vs throughput)
plus TNMMAP (CSS) at d=3/5/7.  Timing: CUDA events around the device-pointer ABI call, inputs resident, 3 warm-ups.
Logical-error counters come from the fused pipeline (`tqec_mc_run`) and are compared with the CPU port on a subsample.
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tensorqec.jl_b200 as tq            # noqa: E402
from tensorqec.jl_b200 import _cabi      # noqa: E402
from oracle import cref, gf2            # noqa: E402


def time_map(plan, words, reps=5):
    B = words.shape[0]
    d_syn = torch.from_numpy(words.view(np.int64)).cuda()
    d_cor = torch.empty((B, plan.ncw), dtype=torch.int64, device="cuda")
    d_lp = torch.empty((B,), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream()
    for _ in range(3):
        plan.decode_map_dev(d_syn.data_ptr(), B, d_cor.data_ptr(), d_lp.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        plan.decode_map_dev(d_syn.data_ptr(), B, d_cor.data_ptr(), d_lp.data_ptr(), st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def time_marginal(plan, words, reps=5):
    B = words.shape[0]
    d_syn = torch.from_numpy(words.view(np.int64)).cuda()
    d_mar = torch.empty((B, 1 << plan.n_obs), dtype=torch.float64, device="cuda")
    d_arg = torch.empty((B,), dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream()
    for _ in range(3):
        plan.decode_marginal_dev(d_syn.data_ptr(), B, d_mar.data_ptr(), d_arg.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        plan.decode_marginal_dev(d_syn.data_ptr(), B, d_mar.data_ptr(), d_arg.data_ptr(), st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


PEAK = {}


def roofline(sch, geom, ms, B):
    """Executed FP64 operations of one launch against the FP64 pipe rate measured on this device: max-plus = one DADD per
    candidate + one DSETP per extra candidate vs the DADD instruction rate; sum-product = one multiply per candidate +
    one add per extra candidate vs the DFMA rate (2 flops per instruction).  Tabulated head steps are not executed."""
    if not PEAK:
        PEAK.update(_cabi.fp64_peak(0))
    if hasattr(sch, "passes"):                                  # global-memory executor: HBM-bound by construction
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6550.0
        ach = sch.bytes_per_shot * B / (ms * 1e-3) / 1e9
        flops = sch.cost * 2 * B / (ms * 1e-3) / 1e12
        return {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "kernel": "k_wide_pass",
                "bytes_per_shot": sch.bytes_per_shot, "fp64_tflops": flops, "fp64_frac_of_dfma": flops / PEAK["dfma_tflops"]}
    if geom.get("table"):
        return None
    h0 = sch.sweep.head_steps if (getattr(sch, "sweep", None) is not None and geom.get("sweep")) else 0
    mul = sum((1 << st.w_out) * len(st.ker) for st in sch.steps[h0:])
    add = sum((1 << st.w_out) * (len(st.ker) - 1) for st in sch.steps[h0:])
    ach = (mul + add) * B / (ms * 1e-3) / 1e12
    if sch.semiring == 0:
        return {"bound": "fp64", "achieved": ach, "peak": PEAK["dadd_tops"], "unit": "TFLOP/s", "frac": ach / PEAK["dadd_tops"],
                "kernel": "k_sweep<maxplus>" if geom.get("sweep") else "k_frontier<maxplus>", "ops_per_shot": mul + add}
    return {"bound": "fp64", "achieved": ach, "peak": PEAK["dfma_tflops"], "unit": "TFLOP/s", "frac": ach / PEAK["dfma_tflops"],
            "kernel": "k_sweep<sumprod>" if geom.get("sweep") else "k_frontier<sumprod>", "ops_per_shot": mul + add}


def css_case(name, code, p, B, seed, out):
    t = tq.CSSTannerGraph(code)
    em = tq.iid_error(p, t)
    mc = tq.MonteCarlo(t, tq.TNMAP(), em)
    plan = mc.plan
    err = _cabi.sample_errors(_cabi.MODEL_DEPOL, [em.px, em.py, em.pz], seed, 0, B)
    syn = mc.H.apply(err)
    ms = time_map(plan, syn)
    counts, mc_ms = mc.run(B, seed=seed)
    # CPU port on a subsample: identical corrections => identical counters
    n = min(B, 20000)
    sch = mc.compiled.cd.schedule if hasattr(mc.compiled, "cd") else mc.compiled.schedule
    bits = tq.unpack_bits(syn[:n], sch.n_checks)
    t0 = time.perf_counter()
    _, cfg = cref.FrontierPlan(sch).run(bits, len(os.sched_getaffinity(0)))
    cpu_rate = n / (time.perf_counter() - t0)
    corr, _ = plan.decode_map(syn[:n], want_logp=False)
    same = bool(np.array_equal(tq.unpack_bits(corr, sch.n_vars), cfg))
    rec = {"case": name, "p": p, "shots": B, "ms": ms, "syndromes_per_s": B / (ms * 1e-3), "pipeline_ms": mc_ms,
           "pipeline_shots_per_s": B / (mc_ms * 1e-3), "logical": {"x": int(counts[0]), "z": int(counts[1]), "any": int(counts[2])},
           "ler": counts[2] / B, "cpu_port_syndromes_per_s": cpu_rate, "cpu_threads": len(os.sched_getaffinity(0)),
           "gpu_equals_cpu_port": same, "geometry": plan.geometry(), "w_max": sch.w_max, "candidates_per_shot": sch.cost,
           "roofline": roofline(sch, plan.geometry(), ms, B)}
    print(json.dumps(rec), file=out, flush=True)
    return rec


def main():
    out = open(os.path.join(ROOT, "gpurun_out", "configs.jsonl"), "w") if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else sys.stdout
    css_case("config1 d=3 TNMAP", tq.SurfaceCode(3, 3), 0.05, 1000, 0, out)
    for d in (5, 7):
        for i, p in enumerate(np.round(np.arange(0.01, 0.1001, 0.01), 2)):
            css_case(f"config2 d={d} TNMAP sweep", tq.SurfaceCode(d, d), float(p), 1_000_000, 1000 * d + i, out)
    for name, code in (("Color488(5)", tq.Color488(5)), ("Steane", tq.SteaneCode())):
        for B in (1, 10, 100, 1000, 10_000, 100_000, 1_000_000):
            css_case(f"config5 {name} TNMAP batch sweep", code, 0.05, B, 5, out)
    # TNMMAP, CSS
    for d, B in ((3, 1_000_000), (5, 1_000_000), (7, 1_000_000), (9, 400_000)):
        t = tq.CSSTannerGraph(tq.SurfaceCode(d, d))
        em = tq.iid_error(0.05, t)
        ct = tq.compile(tq.TNMMAP(), t, em)
        err = _cabi.sample_errors(_cabi.MODEL_DEPOL, [em.px, em.py, em.pz], 77, 0, B)
        from tensorqec.jl_b200.threshold import css_general_matrices
        H, _, _ = css_general_matrices(t, ct.lx, ct.lz)
        syn = _cabi.GF2Matrix(H).apply(err)
        ms = time_marginal(ct.plan, syn)
        print(json.dumps({"case": f"TNMMAP d={d} CSS", "shots": B, "ms": ms, "syndromes_per_s": B / (ms * 1e-3),
                          "geometry": ct.plan.geometry(), "w_max": ct.schedule.w_max, "candidates_per_shot": ct.schedule.cost,
                          "roofline": roofline(ct.schedule, ct.plan.geometry(), ms, B)}),
              file=out, flush=True)
    # DEM TNMMAP (config 4 input format): the reference's fixture + synthetic surface-memory DEMs (benchmarks/make_dem.py)
    for fname, label, B in (("dem.dem", "reference DEM fixture (21 mechanisms, 6 detectors)", 1_000_000),
                            ("surface_d3_r3_phenom.dem", "surface memory d=3 x 3 rounds, phenomenological", 1_000_000),
                            ("generated:3:3", "surface memory d=3 x 3 rounds, circuit-level noise p=1e-3 (circuit.py)", 100_000),
                            ("surface_d5_r5_phenom.dem", "surface memory d=5 x 5 rounds, phenomenological", 200_000),
                            ("generated:5:5", "surface memory d=5 x 5 rounds, circuit-level noise p=1e-3 (circuit.py): 1605 "
                                              "mechanisms, 120 detectors, 29-bit frontier, global-memory executor", 8)):
        if fname.startswith("generated:"):
            _, dd, rr = fname.split(":")
            dem = tq.detector_error_model(tq.parse_stim_string(tq.surface_memory_circuit(
                int(dd), int(rr), "Z", after_clifford_depolarization=1e-3, before_round_data_depolarization=1e-3,
                before_measure_flip_probability=1e-3, after_reset_flip_probability=1e-3)))
        else:
            dem = tq.parse_dem_file(os.path.join(ROOT, "tests", "golden", fname))
        ct = tq.compile(tq.TNMMAP(), dem)
        ep = _cabi.sample_errors(_cabi.MODEL_FLIP, [np.asarray(dem.error_rates)], 3, 0, B)
        syn = _cabi.GF2Matrix(ct.tanner.H).apply(ep)
        ms = time_marginal(ct.plan, syn, reps=2 if B < 100 else 5)
        sch = ct.schedule
        print(json.dumps({"case": f"config4 DEM TNMMAP: {label}", "shots": B, "ms": ms, "syndromes_per_s": B / (ms * 1e-3),
                          "geometry": ct.plan.geometry(), "w_max": getattr(sch, "w_max", None) or getattr(sch, "w_cap", None),
                          "candidates_per_shot": sch.cost, "roofline": roofline(sch, ct.plan.geometry(), ms, B)}),
              file=out, flush=True)
        if fname.startswith("generated:") and dem.n_detectors <= 24:
            # the same problem fully tabulated (opt-in: 2^24 detector patterns decoded once at compile time)
            t0 = time.perf_counter()
            ct2 = tq.compile(tq.TNMMAP(table_bits=24), dem)
            build_s = time.perf_counter() - t0
            B2 = 1_000_000
            ep2 = _cabi.sample_errors(_cabi.MODEL_FLIP, [np.asarray(dem.error_rates)], 3, 0, B2)
            syn2 = _cabi.GF2Matrix(ct2.tanner.H).apply(ep2)
            ms2 = time_marginal(ct2.plan, syn2)
            m1, a1 = ct.plan.decode_marginal(syn2[:2000])
            m2, a2 = ct2.plan.decode_marginal(syn2[:2000])
            print(json.dumps({"case": f"config4 DEM TNMMAP: {label}, TNMMAP(table_bits=24)", "shots": B2, "ms": ms2,
                              "syndromes_per_s": B2 / (ms2 * 1e-3), "compile_s": build_s, "geometry": ct2.plan.geometry(),
                              "table_equals_kernels": bool(np.array_equal(m1, m2) and np.array_equal(a1, a2))}),
                  file=out, flush=True)


if __name__ == "__main__":
    main()
