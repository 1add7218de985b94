#!/bin/bash
# round 2, second session: whole suite, smoke, bench, VAE A/B of the GroupNorm statistics threshold, ncu evidence
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02b_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r02b_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02b_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02b_smoke.log
python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"; python - <<'P'
import json
d=json.loads(open('gpurun_out/r02b_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'unet', d['unet_step_ms'], 'roofline', d['roofline']['frac'], d['roofline']['kernels'], 'gpu_ref', d.get('gpu_reference',{}).get('value'), d['unet_breakdown'])
P
run() { name=$1; shift; env "$@" python scripts/profile_unet.py --vae --out gpurun_out/vae_$name.json > gpurun_out/vae_$name.txt 2>&1; echo "vae512 $name: $(tail -1 gpurun_out/vae_$name.txt)"; }
run cs16384 LDX_GN_STATS_MIN_HW=16384
run cs1024 LDX_GN_STATS_MIN_HW=1024
run cs16384b LDX_GN_STATS_MIN_HW=16384
run cs1024b LDX_GN_STATS_MIN_HW=1024
run2() { name=$1; shift; env "$@" python scripts/profile_unet.py --out gpurun_out/unet_$name.json > gpurun_out/unet_$name.txt 2>&1; echo "unet $name: $(tail -1 gpurun_out/unet_$name.txt)"; }
run2 cs16384 LDX_GN_STATS_MIN_HW=16384
run2 cs1024 LDX_GN_STATS_MIN_HW=1024
run2 cs16384b LDX_GN_STATS_MIN_HW=16384
run2 cs1024b LDX_GN_STATS_MIN_HW=1024
