#!/bin/bash
# A/B of the round-2 kernel changes under bench conditions (sustained load, power cap), same box, alternating
mkdir -p gpurun_out
run() { # name, env...
  local name=$1; shift
  env "$@" python bench.py --no-cpu-baseline --no-gpu-reference --steps 5 --warmup 3 > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - "$name" <<'P'
import json,sys
d=json.loads(open(f'gpurun_out/ab_{sys.argv[1]}.json').read().strip().splitlines()[-1])
b=d['unet_breakdown']
print(f"{sys.argv[1]:14s} value {d['value']:.3f} img/s  step {d['ms_per_step']:.1f} ms  unet {d['unet_step_ms']:.3f} ms  gemm {b['gemm']['ms']:.3f} gn {b['group_norm']['ms']:.3f} attn {b['attention']['ms']:.3f}  clk {d['clocks']['sm_mhz']} W {d['clocks'].get('power_w_max')}")
P
}
run base1 LDX_GEMM_DUAL=0 LDX_CONV_SMALL=0 LDX_FUSE_GN_STATS=0
run new1  LDX_GEMM_DUAL=1
run base2 LDX_GEMM_DUAL=0 LDX_CONV_SMALL=0 LDX_FUSE_GN_STATS=0
run new2  LDX_GEMM_DUAL=1
run dualonly LDX_GEMM_DUAL=1 LDX_CONV_SMALL=0 LDX_FUSE_GN_STATS=0
