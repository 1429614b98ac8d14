# ncu full capture of the sweep kernel for a given library build: gpu_ncu_lib.sh LIB TAG
mkdir -p gpurun_out
TQEC_CUDA_LIB=$PWD/$1 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 2 -c 1 -o gpurun_out/sweep_full_$2 -f python bench.py --steps 2 --warmup 1 --shots 6e5 --cpu-shots 4096 > gpurun_out/ncu_full_$2.log 2>&1
tail -2 gpurun_out/ncu_full_$2.log
