for mt in 512 640; do for sync in 1 2; do
  echo "== maxt=$mt sync=$sync"
  TQEC_SWEEP_MAXT=$mt TQEC_SWEEP_SYNC=$sync python bench.py --steps 3 --warmup 3 --cpu-shots 4096 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print(round(d['value']/1e6,2),'M/s e2e', round(d['e2e']['value']/1e6,2), d['config']['launch']['teams_per_sm'], d['logical_errors']['any'])"
done; done
