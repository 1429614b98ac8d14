# A/B: one CTA of 16 teams per SM (shipped) against two CTAs of 8 teams (TQEC_SWEEP_MAXT=256), same box
mkdir -p gpurun_out
for v in 512 256 512 256; do
  BENCH_NO_ABLATION=1 TQEC_SWEEP_MAXT=$v python bench.py --steps 5 --warmup 3 --cpu-shots 2048 2>gpurun_out/r3b_$v.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print('maxt $v', round(d['value']/1e6,2),'M/s e2e', round(d['e2e']['value']/1e6,2), 'mc', round(d['mc_e2e']['value']/1e6,2), d['logical_errors']['any'], d['e2e']['matches_resident_path'], d['config']['launch'])"
done
tail -3 gpurun_out/r3b_256.err
