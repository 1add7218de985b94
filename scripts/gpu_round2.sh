#!/bin/bash
# kernel tests for the new paths, then A/B per-op tables
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -m gpu -x -q > gpurun_out/r02_kernel_tests.log 2>&1; echo "kernel tests rc=$?"; tail -15 gpurun_out/r02_kernel_tests.log
python -m pytest tests/test_dropin_gpu.py tests/test_engine_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q -s > gpurun_out/r02_model_tests.log 2>&1; echo "model tests rc=$?"; tail -8 gpurun_out/r02_model_tests.log
LDX_GEMM_DUAL=0 LDX_CONV_SMALL=0 python scripts/profile_unet.py --out gpurun_out/r02_ops_unet_base.json > gpurun_out/r02_ops_unet_base.txt 2>&1; tail -1 gpurun_out/r02_ops_unet_base.txt
python scripts/profile_unet.py --out gpurun_out/r02_ops_unet_dual.json > gpurun_out/r02_ops_unet_dual.txt 2>&1; tail -1 gpurun_out/r02_ops_unet_dual.txt
LDX_CONV_SMALL=0 python scripts/profile_unet.py --vae --out gpurun_out/r02_ops_vae512_base.json > gpurun_out/r02_ops_vae512_base.txt 2>&1; tail -1 gpurun_out/r02_ops_vae512_base.txt
python scripts/profile_unet.py --vae --out gpurun_out/r02_ops_vae512_new.json > gpurun_out/r02_ops_vae512_new.txt 2>&1; tail -1 gpurun_out/r02_ops_vae512_new.txt
python scripts/profile_unet.py --vae --batch 4 --latent 128 --out gpurun_out/r02_ops_vae1024_new.json > gpurun_out/r02_ops_vae1024_new.txt 2>&1; tail -1 gpurun_out/r02_ops_vae1024_new.txt
