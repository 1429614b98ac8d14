python -m pytest tests -m gpu -x -q -k "sweep or tnmap or property or fused" 2>&1 | tail -2
for s in 1.25e6 1e7; do
BENCH_NO_ABLATION=1 python bench.py --steps 5 --warmup 3 --cpu-shots 2048 --shots $s 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print('$s shots', round(d['value']/1e6,2),'M/s e2e', round(d['e2e']['value']/1e6,2), 'mc', round(d['mc_e2e']['value']/1e6,2), 'api', round(d['api_e2e']['value']/1e6,2), d['e2e']['matches_resident_path'], d['gpu_launches'])"
done
