# quick GPU check of the sweep kernel: parity tests + headline bench at the three register budgets
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
for mt in 640 768 512; do
  echo "== MAXT $mt"
  TQEC_SWEEP_MAXT=$mt timeout 600 python bench.py --steps 3 --warmup 3 --cpu-shots 4096 2> gpurun_out/bench_$mt.err | tee gpurun_out/bench_$mt.json | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    print(round(d['value']/1e6,2),'M/s e2e',round(d['e2e']['value']/1e6,2), d['config']['launch'], d['logical_errors'], d['e2e']['matches_resident_path'])"
  tail -3 gpurun_out/bench_$mt.err
done
