for c in 21 20 19 18; do
  BENCH_NO_ABLATION=1 TQEC_PIPE_CHUNK_LOG2=$c python bench.py --steps 5 --warmup 3 --cpu-shots 2048 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print('chunk 2^$c', round(d['value']/1e6,2),'M/s e2e', round(d['e2e']['value']/1e6,2), 'mc', round(d['mc_e2e']['value']/1e6,2), 'api', round(d['api_e2e']['value']/1e6,2), d['e2e']['matches_resident_path'])"
done
for c in 19 18 17; do echo "api chunk 2^$c"; TQEC_PIPE_CHUNK_LOG2=$c python benchmarks/api_profile.py 2>&1 | grep -E "total|decode_map_bits"; done
