# same-box A/B of the shipped library against a build variant: ab_variant.sh NAME  (build with build_variant.sh first)
V=tensorqec.jl_b200/csrc/build/variants/libtqec_$1.so
for lib in tensorqec.jl_b200/libtqec_cuda.so $V tensorqec.jl_b200/libtqec_cuda.so $V; do
  BENCH_NO_ABLATION=1 TQEC_CUDA_LIB=$lib python bench.py --steps 5 --warmup 3 --cpu-shots 2048 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print('lib=[$lib]', round(d['value']/1e6,2),'M/s e2e', round(d['e2e']['value']/1e6,2), 'mc', round(d['mc_e2e']['value']/1e6,2), d['logical_errors']['any'], d['e2e']['matches_resident_path'])"
done
