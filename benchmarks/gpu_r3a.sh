# sampler rewrite: parity tests, then the bench (mc_e2e) with the shipped head and with 14 head bits
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "sampler or fused_pipeline or gf2_kernels" > gpurun_out/r3a_pytest.log 2>&1; tail -3 gpurun_out/r3a_pytest.log
python bench.py --steps 5 --warmup 3 --cpu-shots 4096 > gpurun_out/r3a_bench.json 2> gpurun_out/r3a_bench.err
BENCH_NO_ABLATION=1 BENCH_HEAD_BITS=14 python bench.py --steps 5 --warmup 3 --cpu-shots 4096 > gpurun_out/r3a_bench_hb14.json 2> gpurun_out/r3a_bench_hb14.err
python - <<'PY'
import json
for f in ("r3a_bench", "r3a_bench_hb14"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]/1e6, 2), "e2e", round(d["e2e"]["value"]/1e6, 2), "mc", round(d["mc_e2e"]["value"]/1e6, 2),
              "api", round(d["api_e2e"]["value"]/1e6, 2), d["logical_errors"], d["config"]["schedule"]["compile_s"], d["roofline"]["frac"])
    except Exception as e:
        print(f, "ERR", e)
PY
