#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" python scripts/profile_unet.py --out gpurun_out/ab2_$name.json > gpurun_out/ab2_$name.txt 2>&1; echo "$name: $(tail -1 gpurun_out/ab2_$name.txt)"; }
run base LDX_HALO=1
run cs4096 LDX_HALO=1 LDX_GN_STATS_MIN_HW=4096
run cs1024 LDX_HALO=1 LDX_GN_STATS_MIN_HW=1024
run cs256 LDX_HALO=1 LDX_GN_STATS_MIN_HW=256
run h1x1 LDX_HALO=1 LDX_HALO_1X1=1
run base2 LDX_HALO=1
run nohalo LDX_HALO=0
