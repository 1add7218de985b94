#!/bin/bash
# round 2, second session, final state: whole suite, smoke, both bench arms, same-box A/B against the first session's kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02b_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r02b_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02b_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02b_smoke.log
python bench.py > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02b_bench_reference_arm.json 2> gpurun_out/r02b_bench_ref.err; echo "ref rc=$?"
run() { # name, env...
  local name=$1; shift
  env "$@" python bench.py --no-cpu-baseline --no-gpu-reference --steps 5 --warmup 3 > gpurun_out/ab6_$name.json 2> gpurun_out/ab6_$name.err
  python - "$name" <<'P'
import json,sys
d=json.loads(open(f'gpurun_out/ab6_{sys.argv[1]}.json').read().strip().splitlines()[-1])
b=d['unet_breakdown']
print(f"{sys.argv[1]:10s} value {d['value']:.3f} img/s  step {d['ms_per_step']:.1f} ms  unet {d['unet_step_ms']:.3f} ms  gemm+halo {d['roofline']['ms_per_unet_step']:.3f} gn {b['group_norm']['ms']:.3f} attn {b['attention']['ms']:.3f}  clk {d['clocks']['sm_mhz']} W {d['clocks'].get('power_w_max')}")
P
}
run old1 LDX_HALO=0 LDX_XATTN=0 LDX_DIAG64=0
run new1 LDX_HALO=1
run old2 LDX_HALO=0 LDX_XATTN=0 LDX_DIAG64=0
run new2 LDX_HALO=1
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02b_bench_n1.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'unet', d['unet_step_ms'], 'roofline', d['roofline']['frac'], d['roofline']['frac_executed'], d['roofline']['kernels'], 'gpu_ref', d.get('gpu_reference'), d['unet_breakdown'], d['clocks'])
print(open('gpurun_out/r02b_bench_reference_arm.json').read()[-900:])
P
