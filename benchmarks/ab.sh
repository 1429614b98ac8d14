for lib in benchmarks/lib_diag_NOMEM.so benchmarks/lib_diag_NOLAYERS.so; do
  echo "== lib=$lib"
  TQEC_CUDA_LIB=$PWD/$lib python bench.py --steps 3 --warmup 3 --cpu-shots 4096 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print(round(d['value']/1e6,2),'M/s', d['config']['launch']['teams_per_sm'])"
done
