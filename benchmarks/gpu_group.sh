# Experiment: shots per deferred-traceback group of k_sweep (TQEC_SWEEP_GROUP) -> throughput and measured DRAM traffic;
# plus the FP64 issue-model microbenchmark.
mkdir -p gpurun_out
T=${TAG:-r2e}
./benchmarks/micro/issue_model > gpurun_out/${T}_issue_model.jsonl 2>&1; cat gpurun_out/${T}_issue_model.jsonl
for g in ${GROUPS_LIST:-8 32}; do
  TQEC_SWEEP_GROUP=$g timeout 600 python -m pytest tests -m gpu -x -q -k "sweep or d9 or fused_pipeline" 2>&1 | tail -2
done
for g in ${GROUPS_BENCH:-32 16 8 4}; do
  echo "== group $g"
  TQEC_SWEEP_GROUP=$g BENCH_NO_ABLATION=1 timeout 600 python bench.py --steps 3 --warmup 3 --cpu-shots 4096 2> gpurun_out/bench.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    print(round(d['value']/1e6,2),'M/s e2e',round(d['e2e']['value']/1e6,2), 'frac', round(d['roofline']['frac'],4), 'traffic GB/launch', d['roofline']['traffic'] and round(d['roofline']['traffic']/1e9,2), 'LER', d['logical_errors']['any'], d['e2e']['matches_resident_path'])"
  tail -2 gpurun_out/bench.err
done
