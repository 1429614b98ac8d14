#!/usr/bin/env python
"""bench.py -- syndromes decoded per second, TNMAP, d=9 rotated surface code, depolarizing p=0.05 (BASELINE configs[2]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--shots S]

One STEP = one pass of the decode hot path over one batch of S syndromes per GPU (default 1e7, the BASELINE batch).
  value   : whole-job syndromes/s with the bit-packed syndromes already resident in HBM (one kernel launch per step,
            timed with CUDA events on the launching stream, max over ranks).
  e2e     : the same metric through the C-ABI call a Julia / Python host makes (`tqec_decode_map`) with HOST buffers
            (pinned): H2D of the syndromes + kernel + D2H of corrections and log-weights inside the timed region.
  roofline: the decode kernel against the FP64 CUDA-core pipe (max-plus = one DADD + one DSETP per candidate); the
            denominator is measured in this run by `tqec_fp64_peak` because MEASURED_PEAKS.json has no FP64 entry;
            the HBM view (algorithmic bytes per shot vs the measured copy bandwidth) is reported beside it.
  cpu_baseline: the C port of the same frontier recurrence (oracle/csrc/oracle.c) on all host cores, bounded sample.
`--impl reference` times the reference's algorithm on the host cores: pairwise contraction of the DENSE network
(unity vectors, dense parity tensors, greedy tree) one shot at a time, OpenMP over shots -- a compiled, optimistic
stand-in for the Julia reference, which cannot run in this image (no Julia toolchain).
Multi-GPU: one process per GPU (torchrun), shots sharded as contiguous global ranges, no data-path collective; the
only exchange is the all-reduce of the logical-error counters of the fused Monte-Carlo pipeline (outside the timing).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D = 9
HEAD_BITS = 14          # compile-time knob of the sweep lowering: the first 19 steps are tabulated (67 MB + 201 MB tables, 4 s)
P_ERR = 0.05
METRIC = "syndromes decoded/sec (TNMAP, d=9 surface code)"
UNIT = "syndromes/s"


def host_threads():
    """All host cores this process may use (torchrun exports OMP_NUM_THREADS=1; the CPU arms set their thread count
    explicitly instead)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def env_int(name, default):
    return int(os.environ.get(name, default))


def workload_name(shots):
    return (f"d={D} rotated surface code, TNMAP, code-capacity depolarizing p={P_ERR}, {shots:.0e} syndromes per GPU "
            f"per step, contiguous global shot ranges per rank (BASELINE configs[2])")


# ---------------------------------------------------------------------------------------------------------------------
def run_reference(args):
    """Reference arm: the dense-network contraction (what OMEinsum / TensorInference execute per decode call)."""
    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    if rank != 0:
        return 0
    os.environ["TQEC_NO_SWEEP"] = "1"       # the CPU arms only need the plain schedule (skips the head tabulation)
    import tensorqec.jl_b200 as tq          # host data model only (codes, Tanner graph); no GPU call on this path
    from oracle import cref, gf2, networks, philox
    t = tq.CSSTannerGraph(tq.SurfaceCode(D, D))
    em = tq.iid_error(P_ERR, t)
    nq, s2q, pix, pri = networks.general_problem_css(t, em.px, em.py, em.pz)
    dp = cref.DensePlan(networks.tnmap_network(nq, s2q, pix, pri), len(s2q), nq, True)
    threads = host_threads()
    # calibrate the bounded sample: ~4 s of wall time per step
    ex, ez = philox.sample_depolarizing(em.px, em.py, em.pz, 9, 0, 4096)
    sx, sz = gf2.css_syndrome(ex, ez, t.stgx.H, t.stgz.H)
    syn = np.concatenate([sx, sz], axis=1)
    t0 = time.perf_counter()
    dp.run(syn[:threads], threads)
    per_round = time.perf_counter() - t0
    n = args.cpu_shots or int(min(4096, max(threads, threads * max(1, int(4.0 / max(per_round, 1e-3))))))
    for _ in range(args.warmup):
        dp.run(syn[:n], threads)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        lp, cfg = dp.run(syn[:n], threads)
        times.append(time.perf_counter() - t0)
    total = sum(times)
    value = n * args.steps / total
    # the frontier port on the same sample, for context
    _, _, sch = _frontier_schedule(tq)
    fp = cref.FrontierPlan(sch)
    t0 = time.perf_counter()
    fp.run(syn[:n], threads)
    port = n / (time.perf_counter() - t0)
    sample = f"{n} syndromes per step (Philox seed 9, shots 0..{n - 1}), dense greedy tree sc=16, {dp.ops_per_shot:.3g} candidate ops per shot, forward + traceback"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.shots), "arm": "reference algorithm restated in C (oracle/csrc/oracle.c: "
                   "dense pairwise contraction, one shot at a time, OpenMP over shots) on the host cores; the Julia "
                   "reference cannot run in this image", "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "frontier_port_value": port, "host_cpus": os.cpu_count(),
    }
    print(json.dumps(line))
    return 0


def _frontier_schedule(tq):
    t = tq.CSSTannerGraph(tq.SurfaceCode(D, D))
    em = tq.iid_error(P_ERR, t)
    gdp, _ = tq.reduce2general(t, em)
    return t, em, tq.tnmap_schedule(tq.TNMAP(optimizer=_order(), head_bits=HEAD_BITS), gdp)


def _order():
    """BENCH_ORDER=boustro selects the boustrophedon sweep (experiments); default = the planner's choice."""
    if os.environ.get("BENCH_ORDER") == "boustro":
        return [i * D + (j if i % 2 == 0 else D - 1 - j) for i in range(D) for j in range(D)]
    return None


# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is not None:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for ln in self.f.read().splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import tensorqec.jl_b200 as tq
    from tensorqec.jl_b200 import _cabi, sharding

    B = int(args.shots)
    t, em, _ = _frontier_schedule(tq)
    mc = tq.MonteCarlo(t, tq.TNMAP(optimizer=_order(), device=local, head_bits=HEAD_BITS), em)
    plan = mc.plan
    sch = plan.sch
    geom = plan.geometry()
    nsw, ncw = plan.nsw, plan.ncw

    # synthetic input: Philox errors -> syndromes for this rank's global shot range, produced by the library itself
    lo = rank * B
    err_words = _cabi.sample_errors(_cabi.MODEL_DEPOL, [em.px, em.py, em.pz], 9, lo, B, local)
    syn_words = mc.H.apply(err_words)
    h_syn = torch.from_numpy(syn_words.view(np.int64)).pin_memory()
    h_cor = torch.empty((B, ncw), dtype=torch.int64).pin_memory()
    h_lp = torch.empty((B,), dtype=torch.float64).pin_memory()
    d_syn = h_syn.cuda(non_blocking=False)
    d_cor = torch.empty((B, ncw), dtype=torch.int64, device="cuda")
    d_lp = torch.empty((B,), dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream()

    def step_resident():
        plan.decode_map_dev(d_syn.data_ptr(), B, d_cor.data_ptr(), d_lp.data_ptr(), stream.cuda_stream)

    def step_e2e():
        _cabi.check(_cabi.lib().tqec_decode_map(plan.h, h_syn.data_ptr(), B, h_cor.data_ptr(), h_lp.data_ptr()))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        tt = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    # ---- value: resident inputs, device-timed --------------------------------------------------------------------
    for _ in range(args.warmup):
        step_resident()
    launches0 = plan.query(_cabi.Q_LAUNCHES)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_resident()
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = plan.query(_cabi.Q_LAUNCHES) - launches0
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---- e2e: host buffers through the C-ABI call ------------------------------------------------------------------
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = world * B * args.steps / e2e_s
    # results of both paths agree (same kernel, same inputs)
    same = bool(torch.equal(h_cor, d_cor.cpu()) and torch.equal(h_lp, d_lp.cpu()))

    # ---- ablation: the same decode with the shortest tabulated head (6 syndrome bits, 11 of the 81 steps looked up) ------
    ablation = None
    if rank == 0 and geom.get("sweep"):
        gdp_a, _ = tq.reduce2general(t, em)
        sch_a = tq.tnmap_schedule(tq.TNMAP(optimizer=_order(), device=local, head_bits=6), gdp_a)
        plan_a = _cabi.Plan(sch_a, local)
        d_cor_a = torch.empty_like(d_cor)
        d_lp_a = torch.empty_like(d_lp)
        for _ in range(2):
            plan_a.decode_map_dev(d_syn.data_ptr(), B, d_cor_a.data_ptr(), d_lp_a.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        for _ in range(2):
            plan_a.decode_map_dev(d_syn.data_ptr(), B, d_cor_a.data_ptr(), d_lp_a.data_ptr(), stream.cuda_stream)
        a1.record(stream)
        torch.cuda.synchronize()
        ablation = {"head_bits": 6, "tabulated_head_steps": sch_a.sweep.head_steps,
                    "value_one_gpu": B * 2 / (a0.elapsed_time(a1) * 1e-3), "unit": UNIT,
                    "identical_results": bool(torch.equal(d_cor_a, d_cor) and torch.equal(d_lp_a, d_lp))}
        plan_a.close()
        del d_cor_a, d_lp_a
    if world > 1:
        dist.barrier()

    # ---- logical error counters through the fused pipeline + the one collective -------------------------------------
    ler_shots = min(B, 1 << 20)
    counts, mc_ms = mc.run(ler_shots, seed=9, shot_offset=rank * ler_shots)
    counts = sharding.allreduce_counts(counts, local)

    line = None
    if rank == 0:
        peak = _cabi.fp64_peak(local)
        mul_all, add_all = sch.ops_per_shot()
        # executed per shot: the steps after the tabulated head (the head's steps are a table look-up, not arithmetic)
        h0 = sch.sweep.head_steps if (getattr(sch, "sweep", None) is not None and geom.get("sweep")) else 0
        mul = sum((1 << st.w_out) * len(st.ker) for st in sch.steps[h0:])
        add = sum((1 << st.w_out) * (len(st.ker) - 1) for st in sch.steps[h0:])
        ops_per_launch = float(mul + add) * B                  # one add per candidate, one compare per extra candidate
        dur_s = ms_total * 1e-3 / max(launches, 1)
        achieved = ops_per_launch / dur_s / 1e12
        peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm_peak, hbm_src = 6650.0, "fallback"
        if os.path.exists(peaks_file):
            try:
                hbm_peak, hbm_src = float(json.load(open(peaks_file))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
            except Exception:
                pass
        bytes_per_shot = 8 * (nsw + ncw + 1)
        hbm_achieved = bytes_per_shot * B / dur_s / 1e9
        # DRAM traffic of the decode kernel per shot from the newest committed ncu capture (back-pointer scratch streaming
        # through HBM; the algorithmic I/O is 48 B per shot)
        traffic_per_shot, traffic_src = _ncu_traffic_per_shot()
        kname = "k_sweep<maxplus>" if geom.get("sweep") else "k_frontier_warp<maxplus>"
        roofline = {"bound": "fp64", "achieved": achieved, "peak": peak["dadd_tops"], "unit": "TFLOP/s",
                    "frac": achieved / peak["dadd_tops"], "traffic": traffic_per_shot * B if traffic_per_shot else None,
                    "kernel": kname, "ops_per_shot": mul + add, "schedule_ops_per_shot": mul_all + add_all,
                    "tabulated_head_steps": h0,
                    "note": "FP64 CUDA-core pipe: one DADD per candidate + one DSETP per extra candidate of the EXECUTED "
                            "(frontier) schedule, tabulated head steps excluded; max-plus has no tensor-core form",
                    "peak_source": "measured in this run by tqec_fp64_peak (register-resident DADD chains = the FP64 pipe's "
                                   "instruction rate; MEASURED_PEAKS.json has no FP64 entry)",
                    "traffic_source": traffic_src,
                    "fp64_peaks": peak,
                    "hbm": {"bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s",
                            "frac": hbm_achieved / hbm_peak, "bytes_per_shot": bytes_per_shot, "peak_source": hbm_src}}
        cpu = cpu_baseline(tq, sch, syn_words, args)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(B), "shots_per_gpu_per_step": B,
                       "l2": f"inputs larger than L2: {B * nsw * 8 / 1e6:.0f} MB of syndromes in, {B * (ncw + 1) * 8 / 1e6:.0f} MB out per step",
                       "schedule": {"steps": len(sch.steps), "w_max": sch.w_max, "candidates_per_shot": sch.cost,
                                    "head_bits": HEAD_BITS, "note": "roofline ops count only the steps executed per shot (the "
                                    "tabulated head is a table look-up)"},
                       "launch": geom},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * nsw * 8, "d2h_bytes_per_step": B * (ncw + 1) * 8,
                    "api": "tqec_decode_map (host pointers, pinned)", "matches_resident_path": same},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "head_ablation": ablation,
            "logical_errors": {"shots": int(counts[3]), "x": int(counts[0]), "z": int(counts[1]), "any": int(counts[2]),
                               "pipeline_ms_rank0": mc_ms},
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def _ncu_traffic_per_shot():
    """(bytes per shot, source) from the newest profiles/*_ncu_sweep_*_summary.csv: dram read + write of one captured
    launch divided by its shots (the header line of the file states the shot count)."""
    import glob
    import re
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_sweep_*_summary.csv")))
    if not files:
        return None, "no ncu capture committed for this kernel"
    txt = open(files[-1]).read()
    m = re.search(r"shots=(\d+)", txt)
    rd = re.search(r"dram__bytes_read.sum,Gbyte,([0-9.]+)", txt)
    wr = re.search(r"dram__bytes_write.sum,Gbyte,([0-9.]+)", txt)
    if not (m and rd and wr):
        return None, f"could not parse {os.path.basename(files[-1])}"
    per = (float(rd.group(1)) + float(wr.group(1))) * 1e9 / int(m.group(1))
    return per, (f"ncu dram__bytes_read.sum + dram__bytes_write.sum per shot (profiles/{os.path.basename(files[-1])}) "
                 f"x shots per launch")


def cpu_baseline(tq, sch, syn_words, args):
    """C port of the frontier recurrence on all host cores, bounded sample of the same syndromes."""
    from oracle import cref
    fp = cref.FrontierPlan(sch)
    threads = host_threads()
    nchk = sch.n_checks
    probe = tq.unpack_bits(syn_words[:4096], nchk)
    t0 = time.perf_counter()
    fp.run(probe, threads)
    rate = 4096 / (time.perf_counter() - t0)
    n = args.cpu_shots or int(min(syn_words.shape[0], max(4096, rate * 10.0)))
    bits = tq.unpack_bits(syn_words[:n], nchk)
    t0 = time.perf_counter()
    fp.run(bits, threads)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"first {n} syndromes of rank 0's batch, frontier recurrence in C + OpenMP ({dt:.1f} s)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shots", type=float, default=1e7, help="syndromes per GPU per step")
    ap.add_argument("--cpu-shots", type=int, default=0, help="override the bounded CPU sample size")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
