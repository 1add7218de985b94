#!/bin/bash
# same-box A/B: SpatialTransformer.norm folded into proj_in's weights per sample (LDX_GN_FOLD=1) vs gn_apply + linear (0):
# kernel test, graph replay, bench conditions
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "folded_into_proj_in" 2>&1 | tail -1
for e in 1 0 1 0; do LDX_GN_FOLD=$e timeout 300 python scripts/profile_unet.py --out gpurun_out/gnf_$e.json > gpurun_out/gnf_$e.txt 2>&1; echo "fold=$e: $(head -1 gpurun_out/gnf_$e.txt) | $(grep "group_norm  None" gpurun_out/gnf_$e.txt) | $(tail -1 gpurun_out/gnf_$e.txt)"; done
run() { # name, env...
  local name=$1; shift
  env "$@" python bench.py --no-cpu-baseline --no-gpu-reference --steps 5 --warmup 3 > gpurun_out/ab10_$name.json 2> gpurun_out/ab10_$name.err
  python - "$name" <<'P'
import json,sys
d=json.loads(open(f'gpurun_out/ab10_{sys.argv[1]}.json').read().strip().splitlines()[-1])
b=d['unet_breakdown']
print(f"{sys.argv[1]:10s} value {d['value']:.3f} img/s  step {d['ms_per_step']:.1f} ms  unet {d['unet_step_ms']:.3f} ms  gemm family {d['roofline']['ms_per_unet_step']:.3f} gn {b['group_norm']['ms']:.3f} attn {b['attention']['ms']:.3f}  clk {d['clocks']['sm_mhz']} W {d['clocks'].get('power_w_max')}")
P
}
run off1 LDX_GN_FOLD=0
run on1 LDX_GN_FOLD=1
run off2 LDX_GN_FOLD=0
run on2 LDX_GN_FOLD=1
