mkdir -p gpurun_out
T=${TAG:-r2x}
BENCH_NO_ABLATION=1 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 2 -c 1 -o gpurun_out/${T}_sweep_maxplus -f python bench.py --steps 2 --warmup 1 --shots 6e5 --cpu-shots 4096 > gpurun_out/${T}_ncu_maxplus.log 2>&1
tail -2 gpurun_out/${T}_ncu_maxplus.log
python benchmarks/ncu_summary.py gpurun_out/${T}_sweep_maxplus.ncu-rep 600000 > gpurun_out/${T}_ncu_sweep_maxplus_summary.csv
cat gpurun_out/${T}_ncu_sweep_maxplus_summary.csv
