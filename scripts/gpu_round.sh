#!/bin/bash
# One gpurun call: GPU tests, smoke, both bench arms.  Logs under gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/r02_gpu_tests.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/r02_smoke.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"
tail -5 gpurun_out/r02_gpu_tests.log; tail -3 gpurun_out/r02_smoke.log; head -c 1500 gpurun_out/r02_bench.json; echo; head -c 800 gpurun_out/r02_bench_ref.json
