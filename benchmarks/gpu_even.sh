mkdir -p gpurun_out
T=${TAG:-r2s}
timeout 1200 python -m pytest tests -m gpu -x -q -k "sweep or d9 or d3_all or small_codes or fused" > gpurun_out/${T}_pytest_sweep.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest_sweep.log
tail -4 gpurun_out/${T}_pytest_sweep.log
python benchmarks/even_quick.py > gpurun_out/${T}_even.jsonl 2>&1; grep case gpurun_out/${T}_even.jsonl | cut -c1-220
BENCH_NO_ABLATION=1 timeout 600 python bench.py --steps 3 --warmup 3 --cpu-shots 4096 2> gpurun_out/bench.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    print(round(d['value']/1e6,2),'M/s e2e',round(d['e2e']['value']/1e6,2),'frac',round(d['roofline']['frac'],4))"
