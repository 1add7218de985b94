#!/bin/bash
# Kernel-variant library that recompiles ONE translation unit: scripts/build_variant_one.sh NAME UNIT "-DFOO=1 ..."
# -> gpurun_variants/libldx_NAME.so (the other objects come from the regular in-tree build; select with LDX_LIB_PATH).
set -e
cd "$(dirname "$0")/../lightdiffusion_b200/csrc"
NAME=$1; UNIT=$2; EXTRA=$3
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr $EXTRA"
mkdir -p build/v_$NAME ../../gpurun_variants
$NVCC $FLAGS -c $UNIT.cu -o build/v_$NAME/$UNIT.o 2> build/v_$NAME/$UNIT.log
OBJS=""
for f in api gemm elementwise attention conv_small conv_halo xattn ff_fused; do
  if [ $f = $UNIT ]; then OBJS="$OBJS build/v_$NAME/$f.o"; else OBJS="$OBJS build/$f.o"; fi
done
$NVCC -shared -o ../../gpurun_variants/libldx_$NAME.so $OBJS -lcudart_static -lpthread -ldl -lrt
echo built gpurun_variants/libldx_$NAME.so
