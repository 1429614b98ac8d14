#!/usr/bin/env python
"""bench.py -- syndromes decoded per second, TNMAP, d=9 rotated surface code, depolarizing p=0.05 (BASELINE configs[2]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--shots S] [--scaling strong|weak]

One STEP = one pass of the decode hot path over the job's batch of S syndromes (default 1e7, the BASELINE batch),
sharded over the N ranks as contiguous global shot ranges (`--scaling strong`, the default: BASELINE configs[2] says
"1e7 syndromes sharded over 1/2/4/8 GPUs"; `--scaling weak` gives every rank S syndromes).  At N = 1 both are the same.
  value   : whole-job syndromes/s with the bit-packed syndromes already resident in HBM (one kernel launch per step and
            rank, timed with CUDA events on the launching stream, max over ranks).
  e2e     : the same metric through the C-ABI call a Julia / Python host makes (`tqec_decode_map`) with HOST buffers
            (pinned): H2D of the syndromes + kernel + D2H of corrections and log-weights inside the timed region.
  mc_e2e  : the fused Monte-Carlo pipeline (`tqec_mc_run`: Philox sampling -> syndrome extraction -> decode -> logical
            check) over the same shot ranges WITH the one collective -- the NCCL all-reduce of the four counters, issued
            by the library on the pipeline's stream -- inside the timed region.
  api_e2e : (rank 0, N = 1) the user-level call `tq.decode(compiled, CSSSyndrome(sx, sz))` with one byte per bit in host
            memory, i.e. including the host-side packing and unpacking around the C-ABI call.
  roofline: the decode kernel against the FP64 CUDA-core pipe (max-plus = one DADD + one DSETP per candidate); the
            denominator is measured in this run by `tqec_fp64_peak` because MEASURED_PEAKS.json has no FP64 entry;
            `traffic` is measured in this run too (one ncu pass over one launch of the same plan, rank 0, N = 1).
  cpu_baseline: the C port of the same frontier recurrence (oracle/csrc/oracle.c) on all host cores, bounded sample.
`--impl reference` times the reference's algorithm on the host cores: pairwise contraction of the DENSE network
(unity vectors, dense parity tensors, greedy tree) one shot at a time, OpenMP over shots -- a compiled, optimistic
stand-in for the Julia reference, which cannot run in this image (no Julia toolchain).
Multi-GPU: one process per GPU (torchrun), no data-path collective; the decoder is compiled with its shipped defaults.
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
HEAD_BITS = int(os.environ["BENCH_HEAD_BITS"]) if os.environ.get("BENCH_HEAD_BITS") else None   # None = TNMAP's shipped default
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


def workload_name(shots, scaling="strong"):
    per = "per step, sharded over the ranks" if scaling == "strong" else "per GPU per step"
    return (f"d={D} rotated surface code, TNMAP, code-capacity depolarizing p={P_ERR}, {shots:.0e} syndromes {per}, "
            f"contiguous global shot ranges per rank (BASELINE configs[2])")


def _tnmap(tq, **kw):
    if HEAD_BITS is not None:
        kw["head_bits"] = HEAD_BITS
    return tq.TNMAP(optimizer=_order(), **kw)


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
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.shots, args.scaling), "arm": "reference algorithm restated in C (oracle/csrc/oracle.c: "
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
    return t, em, tq.tnmap_schedule(_tnmap(tq), gdp)


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

    total = int(args.shots)
    strong = args.scaling == "strong"
    t = tq.CSSTannerGraph(tq.SurfaceCode(D, D))
    em = tq.iid_error(P_ERR, t)
    t_c0 = time.perf_counter()
    mc = tq.MonteCarlo(t, _tnmap(tq, device=local), em)
    compile_s = time.perf_counter() - t_c0
    plan = mc.plan
    geom = plan.geometry()
    nsw, ncw = plan.nsw, plan.ncw
    comm = sharding.library_comm(local)              # NCCL communicator inside libtqec_cuda.so (None at N = 1)

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

    stream = torch.cuda.current_stream()

    def make_inputs(lo, B):
        """Philox errors -> syndromes for the global shot range [lo, lo + B), produced by the library itself."""
        err_words = _cabi.sample_errors(_cabi.MODEL_DEPOL, [em.px, em.py, em.pz], 9, lo, B, local)
        syn_words = mc.H.apply(err_words)
        h_syn = torch.from_numpy(syn_words.view(np.int64)).pin_memory()
        return syn_words, h_syn

    def timed_resident(lo, B, steps, warmup):
        """K timed launches over resident inputs.  When one step's buffers are smaller than the L2 (strong scaling at
        large N) the steps rotate over enough distinct input / output buffers (further shot ranges of the same size)
        that the footprint exceeds 256 MB, so no step finds its inputs in L2."""
        per_step = B * (nsw + ncw + 1) * 8
        n_buf = max(1, min(16, -(-260_000_000 // per_step)))
        bufs = []
        for k in range(n_buf):
            syn_words, h_syn = make_inputs(lo + k * total * max(world, 1), B)
            bufs.append((syn_words, h_syn, h_syn.cuda(non_blocking=False), torch.empty((B, ncw), dtype=torch.int64, device="cuda"),
                         torch.empty((B,), dtype=torch.float64, device="cuda")))

        def step(i):
            _, _, d_syn, d_cor, d_lp = bufs[i % n_buf]
            plan.decode_map_dev(d_syn.data_ptr(), B, d_cor.data_ptr(), d_lp.data_ptr(), stream.cuda_stream)
        for i in range(warmup):
            step(i)
        n0 = plan.query(_cabi.Q_LAUNCHES)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            step(i)
        e1.record(stream)
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        if n_buf > 1:
            step(0)                                   # leave buffer 0 holding the results of ITS inputs for the e2e comparison
            torch.cuda.synchronize()
        return ms, plan.query(_cabi.Q_LAUNCHES) - n0, bufs[0], n_buf

    # ---- value: resident inputs, device-timed ----------------------------------------------------------------------
    lo, hi = sharding.shard_range(total, rank, world) if strong else (rank * total, (rank + 1) * total)
    B = hi - lo
    sampler = ClockSampler(local)
    sampler.start()
    ms_total, launches, (syn_words, h_syn, d_syn, d_cor, d_lp), n_buf = timed_resident(lo, B, args.steps, args.warmup)
    clocks = sampler.stop()
    job_shots = total if strong else total * world
    value = job_shots * args.steps / (ms_total * 1e-3)

    # ---- e2e: host buffers through the C-ABI call --------------------------------------------------------------------
    h_cor = torch.empty((B, ncw), dtype=torch.int64).pin_memory()
    h_lp = torch.empty((B,), dtype=torch.float64).pin_memory()
    step_e2e = lambda: _cabi.check(_cabi.lib().tqec_decode_map(plan.h, h_syn.data_ptr(), B, h_cor.data_ptr(), h_lp.data_ptr()))
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = job_shots * args.steps / e2e_s
    same = bool(torch.equal(h_cor, d_cor.cpu()) and torch.equal(h_lp, d_lp.cpu()))   # both paths: same kernel, same inputs

    # ---- mc_e2e: sample -> syndrome -> decode -> check -> all-reduce, all inside the timed region ----------------------------
    mc_steps = max(2, min(args.steps, 5))
    mc.run(B, seed=9, shot_offset=lo, comm=comm)
    barrier()
    t0 = time.perf_counter()
    mc_dev_ms = 0.0
    for _ in range(mc_steps):
        counts, ms1 = mc.run(B, seed=9, shot_offset=lo, comm=comm)
        mc_dev_ms += ms1
    mc_s = max_over_ranks(time.perf_counter() - t0)
    mc_dev_ms = max_over_ranks(mc_dev_ms)
    barrier()
    if comm is None and world > 1:
        counts = sharding.allreduce_counts(counts, local)
    mc_e2e = {"value": job_shots * mc_steps / mc_s, "unit": UNIT, "device_value": job_shots * mc_steps / (mc_dev_ms * 1e-3),
              "steps": mc_steps, "collective": ("ncclAllReduce of 4 x uint64 inside tqec_mc_run, every step" if comm is not None
                                                else "none (single rank)"),
              "counts": {"shots": int(counts[3]), "x": int(counts[0]), "z": int(counts[1]), "any": int(counts[2])}}

    # ---- weak-scaling leg beside a strong-scaling headline -----------------------------------------------------------------
    weak = None
    if world > 1 and strong:
        del d_syn, d_cor, d_lp
        torch.cuda.empty_cache()
        ms_w, _, keep, _ = timed_resident(rank * total, total, 2, 1)
        weak = {"value": total * world * 2 / (ms_w * 1e-3), "unit": UNIT, "shots_per_gpu_per_step": total, "steps": 2}
        del keep

    # ---- rank 0 extras: user-level API, head ablation, traffic, CPU baseline -----------------------------------------------------
    line = None
    if rank == 0:
        api = None
        if world == 1:
            nb = min(B, 2_000_000)
            bits = tq.unpack_bits(syn_words[:nb], mc.H.rows)
            nsx = t.stgx.ns
            # the reference's CSSSyndrome holds sx and sz as two arrays of their own (src/decoding/interfaces.jl)
            syn_obj = tq.CSSSyndrome(np.ascontiguousarray(bits[:, :nsx]), np.ascontiguousarray(bits[:, nsx:]))
            tq.decode(mc.compiled, syn_obj)
            t0 = time.perf_counter()
            res = tq.decode(mc.compiled, syn_obj)
            dt = time.perf_counter() - t0
            sx_raw, sz_raw = np.asarray(syn_obj.sx), np.asarray(syn_obj.sz)      # plain arrays again: the constructor scans them
            t0 = time.perf_counter()
            tq.decode(mc.compiled, tq.CSSSyndrome(sx_raw, sz_raw))
            dt_ctor = time.perf_counter() - t0
            api = {"value": nb / dt, "value_incl_constructor": nb / dt_ctor, "unit": UNIT, "shots": nb, "api": "tq.decode(compiled, syn) on a CSSSyndrome(sx, sz) built "
                   "before the timing (its constructor checks the bits once, ~5 ns per shot): one byte per bit in pageable host "
                   "memory in and out, bit packing inside the call", "matches": bool(np.array_equal(
                       res.logp, h_lp.numpy()[:nb]))}
        ablation = None
        if world == 1 and geom.get("sweep") and os.environ.get("BENCH_NO_ABLATION") is None:
            mc_a = tq.MonteCarlo(t, tq.TNMAP(optimizer=_order(), device=local, head_bits=6), em)
            d_syn_a = h_syn.cuda()
            d_cor_a = torch.empty((B, ncw), dtype=torch.int64, device="cuda")
            d_lp_a = torch.empty((B,), dtype=torch.float64, device="cuda")
            run_a = lambda: mc_a.plan.decode_map_dev(d_syn_a.data_ptr(), B, d_cor_a.data_ptr(), d_lp_a.data_ptr(), stream.cuda_stream)
            for _ in range(2):
                run_a()
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            for _ in range(2):
                run_a()
            a1.record(stream)
            torch.cuda.synchronize()
            ablation = {"head_bits": 6, "value_one_gpu": B * 2 / (a0.elapsed_time(a1) * 1e-3), "unit": UNIT,
                        "identical_results": bool(torch.equal(d_cor_a.cpu(), h_cor) and torch.equal(d_lp_a.cpu(), h_lp))}
            del d_syn_a, d_cor_a, d_lp_a, mc_a
        sch = mc.compiled.cd.schedule                     # the Python lowering of the same plan (oracle of the library's): op counts
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
        traffic, traffic_src = (None, "skipped (N > 1)") if world > 1 else measure_traffic(B)
        kname = "k_sweep<maxplus>" if geom.get("sweep") else "k_frontier_warp<maxplus>"
        roofline = {"bound": "fp64", "achieved": achieved, "peak": peak["dadd_tops"], "unit": "TFLOP/s",
                    "frac": achieved / peak["dadd_tops"], "traffic": traffic,
                    "kernel": kname, "ops_per_shot": mul + add, "schedule_ops_per_shot": mul_all + add_all,
                    "tabulated_head_steps": h0,
                    "note": "FP64 CUDA-core pipe: one DADD per candidate + one DSETP per extra candidate of the EXECUTED "
                            "(frontier) schedule, tabulated head steps excluded; max-plus has no tensor-core form",
                    "peak_source": "measured in this run by tqec_fp64_peak (register-resident DADD chains = the FP64 pipe's "
                                   "instruction rate; MEASURED_PEAKS.json has no FP64 entry)",
                    "traffic_source": traffic_src,
                    "fp64_peaks": peak,
                    # the same ops against what the device sustains on the max-plus instruction mix itself: register-resident
                    # chains of 2 DADD + DSETP + 2 FSEL (sm_100a has no FP64 max / 64-bit select); FP64, ALU and FMA pipes each
                    # take one warp instruction per two cycles (benchmarks/micro/issue_model.cu), so the selects cap this mix
                    # at ~0.64 of the DADD rate before any load, store or address instruction
                    "semiring": {"peak": peak["maxplus_tops"], "unit": "Tops/s", "frac": achieved / peak["maxplus_tops"],
                                 "peak_source": "measured in this run by tqec_fp64_peak (max-plus candidate chains)"},
                    "hbm": {"bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s",
                            "frac": hbm_achieved / hbm_peak, "bytes_per_shot": bytes_per_shot, "peak_source": hbm_src}}
        cpu = cpu_baseline(tq, sch, syn_words, args)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(total, args.scaling), "shots_per_step": job_shots, "shots_per_gpu_per_step": B,
                       "l2": f"inputs larger than L2: {B * nsw * 8 / 1e6:.0f} MB of syndromes in, {B * (ncw + 1) * 8 / 1e6:.0f} MB out per GPU and step"
                             + (f"; the steps rotate over {n_buf} distinct buffer sets (>= 260 MB footprint)" if n_buf > 1 else ""),
                       "schedule": {"steps": len(sch.steps), "w_max": sch.w_max, "candidates_per_shot": sch.cost,
                                    "head_bits": len(sch.sweep.head_bits) if getattr(sch, "sweep", None) is not None else 0,
                                    "lowering": "tqec_lower (C++, inside libtqec_cuda.so)", "compile_s": compile_s,
                                    "note": "roofline ops count only the steps executed per shot (the tabulated head is a table look-up)"},
                       "launch": geom},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * nsw * 8, "d2h_bytes_per_step": B * (ncw + 1) * 8,
                    "api": "tqec_decode_map (host pointers, pinned)", "matches_resident_path": same},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "mc_e2e": mc_e2e,
            "api_e2e": api,
            "weak": weak,
            "head_ablation": ablation,
            "logical_errors": mc_e2e["counts"],
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        if comm is not None:
            comm.close()
        dist.destroy_process_group()
    return 0


def measure_traffic(B):
    """DRAM bytes of ONE decode launch over B shots of the same plan, measured now: one `ncu` pass (two counters, no
    clock control) around `bench.py --traffic-probe`.  -> (bytes per launch or None, how it was obtained)."""
    import csv
    import shutil
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None, "ncu not found"
    shots = min(B, 2_000_000)
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-k", "regex:k_sweep|k_frontier",
           "-c", "1", "--launch-skip", "2", "--csv", sys.executable, os.path.abspath(__file__), "--traffic-probe", str(shots)]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=240).stdout
    except Exception as e:                                   # noqa: BLE001
        return None, f"ncu failed: {e}"
    tot = 0.0
    found = 0
    for row in csv.reader(out.splitlines()):
        if len(row) > 3 and row[-3].startswith("dram__bytes_"):
            unit, val = row[-2], float(row[-1].replace(",", ""))
            tot += val * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
            found += 1
    if found < 2:
        return None, "ncu printed no DRAM counters (profiling not permitted?)"
    return tot * B / shots, (f"ncu dram__bytes_read.sum + dram__bytes_write.sum of one launch over {shots} shots of this plan, "
                             f"measured in this run and scaled to {B} shots")


def traffic_probe(shots):
    """Child of measure_traffic: four launches of the benchmark plan on `shots` syndromes (ncu profiles the third)."""
    import torch
    import tensorqec.jl_b200 as tq
    from tensorqec.jl_b200 import _cabi
    t = tq.CSSTannerGraph(tq.SurfaceCode(D, D))
    em = tq.iid_error(P_ERR, t)
    mc = tq.MonteCarlo(t, _tnmap(tq), em)
    err = _cabi.sample_errors(_cabi.MODEL_DEPOL, [em.px, em.py, em.pz], 9, 0, shots, 0)
    d_syn = torch.from_numpy(mc.H.apply(err).view(np.int64)).cuda()
    d_cor = torch.empty((shots, mc.plan.ncw), dtype=torch.int64, device="cuda")
    d_lp = torch.empty((shots,), dtype=torch.float64, device="cuda")
    for _ in range(4):
        mc.plan.decode_map_dev(d_syn.data_ptr(), shots, d_cor.data_ptr(), d_lp.data_ptr(), 0)
    torch.cuda.synchronize()
    return 0


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
    ap.add_argument("--shots", type=float, default=1e7, help="syndromes per step (whole job if strong, per GPU if weak)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--traffic-probe", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--cpu-shots", type=int, default=0, help="override the bounded CPU sample size")
    args = ap.parse_args()
    if args.traffic_probe:
        return traffic_probe(args.traffic_probe)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
