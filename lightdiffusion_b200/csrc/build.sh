#!/bin/bash
# Builds libldx.so in-tree for sm_100a (nvcc cross-compiles without a GPU).  Translation units compile in parallel.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr"
OBJS=""
PIDS=""
mkdir -p build
for f in api gemm elementwise attention conv_small conv_halo xattn ff_fused; do
  if [ ! -f build/$f.o ] || [ $f.cu -nt build/$f.o ] || [ common.cuh -nt build/$f.o ] || [ internal.h -nt build/$f.o ] || [ gemm_tile.cuh -nt build/$f.o ] || [ ../../include/ldx.h -nt build/$f.o ]; then
    rm -f build/$f.o
    ( $NVCC $FLAGS -c $f.cu -o build/$f.o 2> build/$f.log || { cat build/$f.log; exit 1; } ) &
    PIDS="$PIDS $!"
  fi
  OBJS="$OBJS build/$f.o"
done
for p in $PIDS; do wait $p || exit 1; done
$NVCC -shared -o ../libldx.so $OBJS -lcudart_static -lpthread -ldl -lrt
echo "built $(cd .. && pwd)/libldx.so"
