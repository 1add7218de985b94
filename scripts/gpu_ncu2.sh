#!/bin/bash
# ncu evidence for profiles/ (round 2, second session); every ncu command follows a plain run of the same command line
mkdir -p gpurun_out
F='regex:ldx_|gn_|splitk_|cfg_|nchw_|timestep_|eps_to'
python scripts/profile_unet.py --eager-only > gpurun_out/ncu_plain_unet.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$F" -c 1600 --csv \
    --log-file gpurun_out/r02b_unet_step_launches.csv python scripts/profile_unet.py --eager-only > gpurun_out/ncu_unet.log 2>&1
echo "launch list rc=$?"
for c in halo_conv320 halo_conv1280 halo_vae128 xattn40 gemm_dual_lin; do
  case $c in
    halo_*) K='regex:ldx_conv_halo_kernel';;
    xattn40) K='regex:ldx_xattn_kernel';;
    *) K='regex:ldx_gemm_kernel';;
  esac
  python scripts/ncu_targets.py $c > gpurun_out/ncu_plain_$c.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k "$K" -s 2 -c 2 -f -o gpurun_out/r02b_ncu_$c python scripts/ncu_targets.py $c > gpurun_out/ncu_$c.log 2>&1
  echo "$c rc=$?"
done
ls -la gpurun_out/*.ncu-rep | tail
