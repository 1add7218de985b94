#!/bin/bash
mkdir -p gpurun_out
run() { # name, env...
  local name=$1; shift
  env "$@" python bench.py --no-cpu-baseline --no-gpu-reference --steps 5 --warmup 3 > gpurun_out/ab7_$name.json 2> gpurun_out/ab7_$name.err
  python - "$name" <<'P'
import json,sys
d=json.loads(open(f'gpurun_out/ab7_{sys.argv[1]}.json').read().strip().splitlines()[-1])
b=d['unet_breakdown']
print(f"{sys.argv[1]:10s} value {d['value']:.3f} img/s  step {d['ms_per_step']:.1f} ms  unet {d['unet_step_ms']:.3f} ms  gemm+halo {d['roofline']['ms_per_unet_step']:.3f} gn {b['group_norm']['ms']:.3f} attn {b['attention']['ms']:.3f}  clk {d['clocks']['sm_mhz']}")
P
}
run cs16384a LDX_GN_STATS_MIN_HW=16384
run cs1024a LDX_GN_STATS_MIN_HW=1024
run cs4096a LDX_GN_STATS_MIN_HW=4096
run cs16384b LDX_GN_STATS_MIN_HW=16384
run cs1024b LDX_GN_STATS_MIN_HW=1024
run cs4096b LDX_GN_STATS_MIN_HW=4096
