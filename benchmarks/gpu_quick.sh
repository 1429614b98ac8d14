# quick GPU check of the sweep kernel: parity tests + headline bench over register budgets / library variants
mkdir -p gpurun_out
if [ -z "$NOTEST" ]; then
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
fi
run() { echo "== $*"; env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --cpu-shots 4096 2> gpurun_out/bench.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    print(round(d['value']/1e6,2),'M/s e2e',round(d['e2e']['value']/1e6,2), 'teams', d['config']['launch']['teams_per_sm'], d['logical_errors']['any'], d['e2e']['matches_resident_path'])"; tail -2 gpurun_out/bench.err; }
for mt in ${MAXTS:-512 576 640 768}; do run TQEC_SWEEP_MAXT=$mt; done
for lib in ${LIBS:-}; do for mt in ${MAXTS:-512 640}; do run TQEC_CUDA_LIB=$PWD/$lib TQEC_SWEEP_MAXT=$mt; done; done
