#!/bin/bash
python -m pytest tests/test_engine_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q -k "esrgan or taesd or ESRGAN or TAESD or upscal" 2>&1 | tail -3
echo "--- implicit GEMM"; LDX_HALO_GENERIC=0 python scripts/bench_esrgan.py 2>&1 | tail -2
echo "--- halo"; LDX_HALO_GENERIC=1 python scripts/bench_esrgan.py 2>&1 | tail -2
