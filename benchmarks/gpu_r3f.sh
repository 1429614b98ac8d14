for c in 6 3 8 12 6; do echo "copy threads $c"; TQEC_COPY_THREADS=$c python benchmarks/api_profile.py 2>&1 | grep -E "total|decode_map_bits"; done
BENCH_NO_ABLATION=1 python bench.py --steps 5 --warmup 3 --cpu-shots 2048 --shots 1.25e6 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print('1.25e6 shots', round(d['value']/1e6,2),'M/s e2e', round(d['e2e']['value']/1e6,2), 'mc', round(d['mc_e2e']['value']/1e6,2), 'api', round(d['api_e2e']['value']/1e6,2), d['e2e']['matches_resident_path'])"
