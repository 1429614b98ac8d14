# one GPU call: parity tests, both bench arms, the ncu launch list, one full capture of the decode kernel, other configs
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
tail -c 1500 gpurun_out/bench_ours.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --cpu-shots 4096 > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 2 -c 1 -o gpurun_out/sweep_full -f python bench.py --steps 2 --warmup 1 --shots 6e5 --cpu-shots 4096 > gpurun_out/ncu_full.log 2>&1
python benchmarks/configs.py > gpurun_out/configs_all.jsonl 2> gpurun_out/configs_all.err
tail -3 gpurun_out/configs_all.err
ls -la gpurun_out
