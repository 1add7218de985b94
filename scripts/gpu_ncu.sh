#!/bin/bash
# ncu evidence for profiles/ (one gpurun call; every ncu command follows a plain run of the same command line)
mkdir -p gpurun_out
F='regex:ldx_|gn_|splitk_|cfg_|nchw_|timestep_|eps_to'
python scripts/profile_unet.py --eager-only > gpurun_out/ncu_plain_unet.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$F" -c 1600 --csv \
    --log-file gpurun_out/r02_unet_step_launches.csv python scripts/profile_unet.py --eager-only > gpurun_out/ncu_unet.log 2>&1
echo "launch list rc=$?"
for c in gemm_dual_conv gemm_dual_conv320 gemm_dual_lin conv_small_vae gn_colstats_vae; do
  case $c in
    conv_small_vae) K='regex:ldx_conv3x3_small';;
    gn_colstats_vae) K='regex:gn_colstats_reduce|gn_apply';;
    *) K='regex:ldx_gemm_kernel';;
  esac
  python scripts/ncu_targets.py $c > gpurun_out/ncu_plain_$c.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k "$K" -s 2 -c 2 -f -o gpurun_out/r02_ncu_$c python scripts/ncu_targets.py $c > gpurun_out/ncu_$c.log 2>&1
  echo "$c rc=$?"
done
ls -la gpurun_out/*.ncu-rep | tail
