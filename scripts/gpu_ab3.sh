#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -m gpu -x -q 2>&1 | tail -3
run2() { name=$1; shift; env "$@" python scripts/profile_unet.py --out gpurun_out/ab3_$name.json > gpurun_out/ab3_$name.txt 2>&1; echo "unet $name: $(tail -1 gpurun_out/ab3_$name.txt)"; }
run2 xa0 LDX_XATTN=0
run2 xa1 LDX_XATTN=1
run2 xa0b LDX_XATTN=0
run2 xa1b LDX_XATTN=1
