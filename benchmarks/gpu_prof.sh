# Profiling call: pipe-model microbenchmark, full ncu captures (with source) of k_sweep<maxplus> and one mid-plan k_wide_pass.
mkdir -p gpurun_out
T=${TAG:-r2d}
./benchmarks/micro/issue_model > gpurun_out/${T}_issue_model.jsonl 2>&1; cat gpurun_out/${T}_issue_model.jsonl
BENCH_NO_ABLATION=1 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 2 -c 1 -o gpurun_out/${T}_sweep_maxplus -f python bench.py --steps 2 --warmup 1 --shots 6e5 --cpu-shots 4096 > gpurun_out/${T}_ncu_maxplus.log 2>&1
tail -2 gpurun_out/${T}_ncu_maxplus.log
ncu --set full --clock-control none --import-source on -k regex:k_wide_pass -s 60 -c 1 -o gpurun_out/${T}_wide_pass -f python benchmarks/wide_d5.py 5 1 > gpurun_out/${T}_ncu_wide.log 2>&1
tail -2 gpurun_out/${T}_ncu_wide.log
ls -la gpurun_out
