#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_engine_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -3
run2() { name=$1; shift; env "$@" python scripts/profile_unet.py --out gpurun_out/ff_$name.json > gpurun_out/ff_$name.txt 2>&1; echo "unet $name: $(tail -1 gpurun_out/ff_$name.txt)"; }
run2 off LDX_FF_FUSED=0
run2 on LDX_FF_FUSED=1
run2 off2 LDX_FF_FUSED=0
run2 on2 LDX_FF_FUSED=1
