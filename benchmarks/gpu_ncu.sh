# ncu full capture of the sweep kernel (one launch, 6e5 shots) at the register budget given by $1
mkdir -p gpurun_out
MT=${1:-512}
TQEC_SWEEP_MAXT=$MT ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 2 -c 1 -o gpurun_out/sweep_full_$MT -f python bench.py --steps 2 --warmup 1 --shots 6e5 --cpu-shots 4096 > gpurun_out/ncu_full_$MT.log 2>&1
tail -3 gpurun_out/ncu_full_$MT.log
