#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py tests/test_conv_halo_gpu.py -m gpu -x -q 2>&1 | tail -4
python -m pytest tests/test_engine_gpu.py -m gpu -x -q 2>&1 | tail -3
run() { name=$1; shift; env "$@" python scripts/profile_unet.py --out gpurun_out/d64_$name.json > gpurun_out/d64_$name.txt 2>&1; echo "$name: $(tail -1 gpurun_out/d64_$name.txt)"; }
run off LDX_DIAG64=0
run on LDX_DIAG64=1
run off2 LDX_DIAG64=0
run on2 LDX_DIAG64=1
run on_cs1024 LDX_DIAG64=1 LDX_GN_STATS_MIN_HW=1024
