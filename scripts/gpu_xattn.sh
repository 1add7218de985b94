#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attention" 2>&1 | tail -8
echo "--- generic kernel"; LDX_XATTN=0 timeout 120 python scripts/bench_cross_attn.py 2>&1 | tail -5
echo "--- xattn kernel"; LDX_XATTN=1 timeout 120 python scripts/bench_cross_attn.py 2>&1 | tail -5
LDX_LIB_PATH=gpurun_variants/libldx_xat.so python scripts/xattn_timeline.py 40; LDX_LIB_PATH=gpurun_variants/libldx_xat.so python scripts/xattn_timeline.py 80
