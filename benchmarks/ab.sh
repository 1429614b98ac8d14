# same-box A/B of two source trees / sync modes on the headline workload
for tree in . ab_late; do
  for sync in 0 1 2; do
    echo "== tree=$tree sync=$sync"
    (cd $tree && TQEC_SWEEP_SYNC=$sync python bench.py --steps 4 --warmup 3 --cpu-shots 4096 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print(round(d['value']/1e6,2),'M/s', d['config']['launch']['teams_per_sm'], d['logical_errors']['any'], d['e2e']['matches_resident_path'])")
  done
done
