# same-box A/B over one environment knob: ab_env.sh NAME VALUE  (unset / set, alternating, bench.py twice each)
for v in "" "$2" "" "$2"; do
  if [ -z "$v" ]; then unset $1; else export $1=$v; fi
  BENCH_NO_ABLATION=1 python bench.py --steps 5 --warmup 3 --cpu-shots 2048 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print('$1=[$v]', round(d['value']/1e6,2),'M/s e2e', round(d['e2e']['value']/1e6,2), 'mc', round(d['mc_e2e']['value']/1e6,2), d['logical_errors']['any'], d['e2e']['matches_resident_path'], d['config']['launch']['smem_bytes'])"
done
