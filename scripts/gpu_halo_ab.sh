#!/bin/bash
# halo-conv integration: model-level parity, then per-op / graph-replay A/B of the UNet step and the VAE decodes
mkdir -p gpurun_out
python -m pytest tests/test_engine_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q > gpurun_out/halo_model_tests.log 2>&1; echo "model tests rc=$?"; tail -4 gpurun_out/halo_model_tests.log
for v in 0 1; do
LDX_HALO=$v python scripts/profile_unet.py --out gpurun_out/halo${v}_ops_unet.json > gpurun_out/halo${v}_ops_unet.txt 2>&1; tail -1 gpurun_out/halo${v}_ops_unet.txt
LDX_HALO=$v python scripts/profile_unet.py --vae --out gpurun_out/halo${v}_ops_vae512.json > gpurun_out/halo${v}_ops_vae512.txt 2>&1; tail -1 gpurun_out/halo${v}_ops_vae512.txt
LDX_HALO=$v python scripts/profile_unet.py --vae --batch 4 --latent 128 --out gpurun_out/halo${v}_ops_vae1024.json > gpurun_out/halo${v}_ops_vae1024.txt 2>&1; tail -1 gpurun_out/halo${v}_ops_vae1024.txt
done
LDX_HALO=1 LDX_HALO_MIN_HW=256 python scripts/profile_unet.py --out gpurun_out/halo256_ops_unet.json > gpurun_out/halo256_ops_unet.txt 2>&1; tail -1 gpurun_out/halo256_ops_unet.txt
LDX_HALO=1 LDX_HALO_MIN_HW=4096 python scripts/profile_unet.py --out gpurun_out/halo4096_ops_unet.json > gpurun_out/halo4096_ops_unet.txt 2>&1; tail -1 gpurun_out/halo4096_ops_unet.txt
