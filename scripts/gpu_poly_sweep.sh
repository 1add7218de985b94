#!/bin/bash
# d = 40 self-attention with every variant library under gpurun_variants/ (packed polynomial exponential masks)
mkdir -p gpurun_out
: > gpurun_out/poly_sweep.txt
for lib in "" gpurun_variants/libldx_*.so ""; do
  if [ -n "$lib" ]; then export LDX_LIB_PATH=$lib; else unset LDX_LIB_PATH; fi
  r=$(timeout 120 python scripts/bench_attn40.py 2>/dev/null | awk '/d=40/ {printf "%s ", $6}')
  echo "${lib:-default} $r" | tee -a gpurun_out/poly_sweep.txt
done
