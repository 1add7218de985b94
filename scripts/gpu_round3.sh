#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -m gpu -x -q > gpurun_out/r02_kernel_tests.log 2>&1; echo "kernel tests rc=$?"; tail -12 gpurun_out/r02_kernel_tests.log
python -m pytest tests/test_dropin_gpu.py tests/test_engine_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q -s > gpurun_out/r02_model_tests.log 2>&1; echo "model tests rc=$?"; tail -6 gpurun_out/r02_model_tests.log
LDX_FUSE_GN_STATS=0 python scripts/profile_unet.py --out gpurun_out/r02_ops_unet_nogn.json > gpurun_out/r02_ops_unet_nogn.txt 2>&1; tail -1 gpurun_out/r02_ops_unet_nogn.txt
python scripts/profile_unet.py --out gpurun_out/r02_ops_unet_gn.json > gpurun_out/r02_ops_unet_gn.txt 2>&1; tail -1 gpurun_out/r02_ops_unet_gn.txt
LDX_FUSE_GN_STATS=0 python scripts/profile_unet.py --vae --out gpurun_out/r02_ops_vae512_nogn.json > gpurun_out/r02_ops_vae512_nogn.txt 2>&1; tail -1 gpurun_out/r02_ops_vae512_nogn.txt
python scripts/profile_unet.py --vae --out gpurun_out/r02_ops_vae512_gn.json > gpurun_out/r02_ops_vae512_gn.txt 2>&1; tail -1 gpurun_out/r02_ops_vae512_gn.txt
python scripts/profile_unet.py --vae --batch 4 --latent 128 --out gpurun_out/r02_ops_vae1024_gn.json > gpurun_out/r02_ops_vae1024_gn.txt 2>&1; tail -1 gpurun_out/r02_ops_vae1024_gn.txt
LDX_LIB_PATH=gpurun_variants/libldx_timing.so python scripts/gemm_timeline.py > gpurun_out/r02_gemm_timeline.txt 2>&1; echo "timeline rc=$?"
