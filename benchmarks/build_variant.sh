#!/bin/bash
# build a variant of libtqec_cuda.so with extra -D flags for tqec_sweep.cu / tqec_wide.cu: build_variant.sh NAME FILE.cu "-DX=1 ..."
# -> tensorqec.jl_b200/csrc/build/variants/libtqec_NAME.so (travels with gpurun, git-ignored); use with TQEC_CUDA_LIB=...
set -e
cd "$(dirname "$0")/../tensorqec.jl_b200/csrc"
NAME=$1; FILE=$2; DEFS=$3
mkdir -p build/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall -I../../include"
nvcc $FLAGS $DEFS -c -o build/variants/${NAME}_${FILE%.cu}.o $FILE
OBJS=""
for o in build/*.o; do if [ "$(basename $o)" == "${FILE%.cu}.o" ]; then OBJS="$OBJS build/variants/${NAME}_${FILE%.cu}.o"; else OBJS="$OBJS $o"; fi; done
nvcc $FLAGS -shared -o build/variants/libtqec_${NAME}.so $OBJS -ldl
echo build/variants/libtqec_${NAME}.so
