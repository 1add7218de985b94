#!/bin/bash
# same-box A/B under bench conditions, third session of round 2: second-session kernels (two-GEMM feed-forward, scalar
# epilogue math: LDX_FF_FUSED=0 + the -DLDX_PACKED_F32=0 build) vs defaults (fused feed-forward, packed fp32 epilogues)
mkdir -p gpurun_out
run() { # name, env...
  local name=$1; shift
  env "$@" python bench.py --no-cpu-baseline --no-gpu-reference --steps 5 --warmup 3 > gpurun_out/ab8_$name.json 2> gpurun_out/ab8_$name.err
  python - "$name" <<'P'
import json,sys
d=json.loads(open(f'gpurun_out/ab8_{sys.argv[1]}.json').read().strip().splitlines()[-1])
b=d['unet_breakdown']
print(f"{sys.argv[1]:10s} value {d['value']:.3f} img/s  step {d['ms_per_step']:.1f} ms  unet {d['unet_step_ms']:.3f} ms  gemm family {d['roofline']['ms_per_unet_step']:.3f} gn {b['group_norm']['ms']:.3f} attn {b['attention']['ms']:.3f}  clk {d['clocks']['sm_mhz']} W {d['clocks'].get('power_w_max')}")
P
}
run old1 LDX_FF_FUSED=0 LDX_LIB_PATH=gpurun_variants/libldx_scalar.so
run new1 LDX_FF_FUSED=1
run old2 LDX_FF_FUSED=0 LDX_LIB_PATH=gpurun_variants/libldx_scalar.so
run new2 LDX_FF_FUSED=1
