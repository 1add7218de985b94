#!/bin/bash
# same-box A/B under bench conditions: packed polynomial exponentials in the d = 40 self-attention off (variant library) / on
mkdir -p gpurun_out
run() { # name, env...
  local name=$1; shift
  env "$@" python bench.py --no-cpu-baseline --no-gpu-reference --steps 5 --warmup 3 > gpurun_out/ab9_$name.json 2> gpurun_out/ab9_$name.err
  python - "$name" <<'P'
import json,sys
d=json.loads(open(f'gpurun_out/ab9_{sys.argv[1]}.json').read().strip().splitlines()[-1])
b=d['unet_breakdown']
print(f"{sys.argv[1]:10s} value {d['value']:.3f} img/s  step {d['ms_per_step']:.1f} ms  unet {d['unet_step_ms']:.3f} ms  gemm family {d['roofline']['ms_per_unet_step']:.3f} gn {b['group_norm']['ms']:.3f} attn {b['attention']['ms']:.3f}  clk {d['clocks']['sm_mhz']} W {d['clocks'].get('power_w_max')}")
P
}
run off1 LDX_LIB_PATH=gpurun_variants/libldx_nopoly.so
run on1 A=1
run off2 LDX_LIB_PATH=gpurun_variants/libldx_nopoly.so
run on2 A=1
