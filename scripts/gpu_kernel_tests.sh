#!/bin/bash
# Runs each kernel test group in its own process (a device trap poisons the CUDA context), logs to gpurun_out/.
mkdir -p gpurun_out
: > gpurun_out/kernel_tests.log
for t in "$@"; do
  echo "=== $t" >> gpurun_out/kernel_tests.log
  timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "$t" -x --no-header -p no:cacheprovider 2>&1 | tail -25 >> gpurun_out/kernel_tests.log
done
grep -E "^===|passed|failed|error|Error|assert" gpurun_out/kernel_tests.log | head -150
