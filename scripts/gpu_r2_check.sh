#!/bin/bash
# round-2 re-entry check: the whole GPU suite, smoke, and the default bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02b_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -5 gpurun_out/r02b_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02b_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02b_smoke.log
python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/r02b_bench.json
python scripts/profile_unet.py --out gpurun_out/r02b_ops_unet.json > gpurun_out/r02b_ops_unet.txt 2>&1; tail -3 gpurun_out/r02b_ops_unet.txt
