# same-box A/B of library builds on the headline workload
for lib in benchmarks/lib_b94c397.so benchmarks/lib_26627cc.so tensorqec.jl_b200/libtqec_cuda.so benchmarks/lib_b94c397.so tensorqec.jl_b200/libtqec_cuda.so; do
  echo "== $lib"
  TQEC_CUDA_LIB=$PWD/$lib python bench.py --steps 4 --warmup 3 --cpu-shots 4096 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print(round(d['value']/1e6,2),'M/s', d['config']['launch']['teams_per_sm'], d['config']['launch']['smem_bytes'])"
done
