#!/bin/bash
# Builds a kernel-variant copy of libldx.so for experiments: scripts/build_variant.sh NAME "-DFOO=1 ..."
# -> gpurun_variants/libldx_NAME.so (select with LDX_LIB_PATH).
set -e
cd "$(dirname "$0")/../lightdiffusion_b200/csrc"
NAME=$1; EXTRA=$2
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr $EXTRA"
mkdir -p build/$NAME ../../gpurun_variants
for f in api gemm elementwise attention conv_small conv_halo xattn ff_fused; do
  $NVCC $FLAGS -c $f.cu -o build/$NAME/$f.o 2> build/$NAME/$f.log &
done
wait
$NVCC -shared -o ../../gpurun_variants/libldx_$NAME.so build/$NAME/*.o -lcudart_static -lpthread -ldl -lrt
echo built gpurun_variants/libldx_$NAME.so
