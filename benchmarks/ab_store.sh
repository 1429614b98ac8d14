for rep in 1 2; do for v in sm0x01 sm0x11 sm0x15 sm0x33; do echo "== $v"; TQEC_CUDA_LIB=$PWD/tensorqec.jl_b200/csrc/build/variants/libtqec_$v.so BENCH_NO_ABLATION=1 python bench.py --steps 3 --warmup 3 --cpu-shots 4096 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(round(d['value']/1e6,2),'M/s', d['logical_errors']['any'])"; done; done
