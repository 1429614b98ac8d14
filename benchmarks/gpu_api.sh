mkdir -p gpurun_out
T=${TAG:-r2l}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest_gpu.log
tail -3 gpurun_out/${T}_pytest_gpu.log
BENCH_NO_ABLATION=1 timeout 600 python bench.py --steps 3 --warmup 3 --cpu-shots 4096 2> gpurun_out/bench.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    print(round(d['value']/1e6,2),'M/s e2e',round(d['e2e']['value']/1e6,2),'api',d['api_e2e'])"
tail -2 gpurun_out/bench.err
