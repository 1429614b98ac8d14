# Round-2 GPU call: parity tests, sanitizer passes on a subset, both bench arms, ncu launch list, full captures of the three
# decode kernels (k_sweep<maxplus>, k_sweep<sumprod>, k_wide_bf), all other configs.  Everything lands in gpurun_out/.
set -x
mkdir -p gpurun_out
T=${TAG:-r2}
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest_gpu.log
tail -4 gpurun_out/${T}_pytest_gpu.log
# compute-sanitizer: memcheck over small-plan tests of every kernel family; racecheck on the shared-memory kernels
SEL='d3_all or small_codes or wide_executor or table_decoder or sector or (sweep_kernel_edge_cases and 7) or gf2_kernels or sampler or dem_tnmmap or tnmmap_golden'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 97 python -m pytest tests -m gpu -q -k "$SEL" > gpurun_out/${T}_sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/${T}_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 97 python -m pytest tests -m gpu -q -k "(sweep_kernel_edge_cases and 7) or wide_executor or small_codes or dem_from_generated or dem_tnmmap" > gpurun_out/${T}_sanitizer_racecheck.log 2>&1; echo "racecheck exit $?" >> gpurun_out/${T}_sanitizer_racecheck.log
tail -3 gpurun_out/${T}_sanitizer_memcheck.log gpurun_out/${T}_sanitizer_racecheck.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench_ours.json 2> gpurun_out/${T}_bench_ours.err
tail -c 600 gpurun_out/${T}_bench_ours.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --cpu-shots 4096 > gpurun_out/${T}_bench_under_ncu.log 2>&1
BENCH_NO_ABLATION=1 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 2 -c 1 -o gpurun_out/${T}_sweep_maxplus -f python bench.py --steps 2 --warmup 1 --shots 6e5 --cpu-shots 4096 > gpurun_out/${T}_ncu_maxplus.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 1 -c 1 -o gpurun_out/${T}_sweep_sumprod -f python benchmarks/tnmmap_quick.py 9 400000 > gpurun_out/${T}_ncu_sumprod.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_wide_bf -s 60 -c 1 -o gpurun_out/${T}_wide_bf -f python benchmarks/wide_d5.py 5 1 > gpurun_out/${T}_ncu_wide.log 2>&1
for k in sweep_maxplus:600000 sweep_sumprod:400000 wide_bf:1; do python benchmarks/ncu_summary.py gpurun_out/${T}_${k%%:*}.ncu-rep ${k##*:} > gpurun_out/${T}_ncu_${k%%:*}_summary.csv 2>gpurun_out/${T}_ncu_${k%%:*}_summary.err; done
python benchmarks/configs.py > gpurun_out/${T}_configs.log 2>&1; cp gpurun_out/configs.jsonl gpurun_out/${T}_configs_all.jsonl
python benchmarks/wide_d5.py 5 8 > gpurun_out/${T}_wide_d5.jsonl 2>&1
python benchmarks/dem_wide_compare.py > gpurun_out/${T}_dem_wide_compare.jsonl 2>&1
ls -la gpurun_out | tail -30
