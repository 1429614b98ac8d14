"""Summarise an ncu report (raw page + SASS source page) of the decode kernel: python benchmarks/ncu_summary.py REP [SHOTS]"""
import csv, subprocess, sys, io
from collections import Counter, defaultdict
rep = sys.argv[1]; shots = float(sys.argv[2]) if len(sys.argv) > 2 else 6e5
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'smsp__warps_active.avg.per_cycle_active', 'smsp__warps_eligible.avg.per_cycle_active', 'sm__cycles_elapsed.avg',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed_pipe_fp64.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active']
print("metric,unit,value")
for w in want:
    if w in m: print(f"{w},{m[w][1]},{m[w][0]}")
if 'smsp__inst_executed.sum' in m:
    print(f"derived__warp_instructions_per_shot,inst,{float(m['smsp__inst_executed.sum'][0]) / shots:.1f}")
for h in hdr:
    if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
        try:
            v = float(m[h][0])
        except ValueError:
            continue
        if v > 0.05: print(f"{h},{m[h][1]},{v:.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h2 = rows[1]; data = rows[2:]
ia, isamp, isrc = h2.index("Instructions Executed"), h2.index("# Samples"), h2.index("Source")
tot = sum(int(r[ia]) for r in data); ts = sum(int(r[isamp]) for r in data)
ops = Counter()
for r in data:
    t = r[isrc].split()
    op = t[1] if t[0].startswith('@') else t[0]
    ops[op.split('.')[0]] += int(r[ia])
print("# instruction mix (share of executed warp instructions)")
for op, n in ops.most_common(14): print(f"mix__{op},%,{100 * n / tot:.1f}")
if len(sys.argv) > 3:
    # hot windows
    W = 64
    for b in range(0, len(data), W):
        s = sum(int(r[isamp]) for r in data[b:b + W]); n = sum(int(r[ia]) for r in data[b:b + W])
        if s > 0.02 * ts: print(f"window {b}: samples {100*s/ts:.1f}% inst {100*n/tot:.1f}%")
