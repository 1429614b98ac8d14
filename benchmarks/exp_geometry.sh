run() { echo "== $*"; env "$@" python bench.py --steps 2 --warmup 2 --cpu-shots 4096 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print(round(d['value']/1e6,2),'M/s', d['config']['launch'], d['config']['schedule'], d['logical_errors'])
"; }
run A=1
run BENCH_ORDER=boustro
run BENCH_ORDER=boustro TQEC_TARGET_BITS=9
run BENCH_ORDER=boustro TQEC_TARGET_BITS=9 TQEC_TEAMS_PER_CTA=20
run BENCH_ORDER=boustro TQEC_TARGET_BITS=9 TQEC_TEAMS_PER_CTA=16
