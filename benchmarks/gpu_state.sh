# State check of HEAD: GPU parity tests, headline bench, then the traceback-group experiment and the issue-model microbenchmark.
mkdir -p gpurun_out
T=${TAG:-r2c}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest_gpu.log
tail -5 gpurun_out/${T}_pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench_ours.json 2> gpurun_out/${T}_bench_ours.err
tail -c 1500 gpurun_out/${T}_bench_ours.json; tail -3 gpurun_out/${T}_bench_ours.err
NOTEST=1 TAG=$T GROUPS_LIST="8" GROUPS_BENCH="32 16 8 4" bash benchmarks/gpu_group.sh
