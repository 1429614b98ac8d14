for mt in 512 640; do echo "== MAXT $mt"; TQEC_SWEEP_MAXT=$mt BENCH_NO_ABLATION=1 python bench.py --steps 3 --warmup 3 --cpu-shots 4096 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(round(d['value']/1e6,2),'M/s', d['config']['launch'])"; done
