for hb in 12 14 16; do
  echo "== head bits=$hb"
  TQEC_HEAD_BITS=$hb python bench.py --steps 3 --warmup 3 --cpu-shots 4096 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print(round(d['value']/1e6,2),'M/s e2e', round(d['e2e']['value']/1e6,2), d['logical_errors']['any'], d['e2e']['matches_resident_path'], d['roofline']['tabulated_head_steps'])"
done
