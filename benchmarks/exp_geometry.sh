# geometry experiments on the headline workload (bench.py honours TQEC_* overrides read by tqec_plan_create)
run() { echo "== $*"; env "$@" python bench.py --steps 2 --warmup 2 --cpu-shots 4096 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print(round(d['value']/1e6,2),'M/s', d['config']['launch'], d['config']['schedule'])
"; }
run A=1
run TQEC_TARGET_BITS=9 TQEC_TEAMS_PER_CTA=22
run TQEC_TARGET_BITS=9 TQEC_TEAMS_PER_CTA=16
run TQEC_TEAMS_PER_CTA=8
