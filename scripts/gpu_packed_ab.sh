#!/bin/bash
# A/B of the packed-fp32 (FFMA2 / FMUL2) epilogues: default build vs gpurun_variants/libldx_scalar.so (-DLDX_PACKED_F32=0 -DLDX_FF_PACKED=0)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q 2>&1 | tail -2
for v in packed scalar packed scalar; do
  if [ $v = scalar ]; then export LDX_LIB_PATH=gpurun_variants/libldx_scalar.so; else unset LDX_LIB_PATH; fi
  echo "== $v"
  timeout 120 python scripts/bench_attn40.py
  timeout 120 python scripts/bench_cross_attn.py 2>&1 | tail -3
  timeout 120 python scripts/bench_ff.py
  timeout 300 python scripts/profile_unet.py --out gpurun_out/pk_$v.json > gpurun_out/pk_$v.txt 2>&1; echo "unet $v: $(tail -1 gpurun_out/pk_$v.txt)"
done
