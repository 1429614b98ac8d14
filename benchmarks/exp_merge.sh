for m in 0 1; do echo "== TQEC_MERGE=$m"; TQEC_MERGE=$m python bench.py --steps 3 --warmup 2 --cpu-shots 4096 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print(round(d['value']/1e6,2),'M/s', d['config']['launch']['teams_per_sm'], d['config']['schedule'], d['logical_errors'])"; done
