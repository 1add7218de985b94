#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -m gpu -x -q > gpurun_out/r02_kernel_tests.log 2>&1; echo "kernel tests rc=$?"; tail -12 gpurun_out/r02_kernel_tests.log
python -m pytest tests/test_engine_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q -s > gpurun_out/r02_model_tests.log 2>&1; echo "model tests rc=$?"; tail -4 gpurun_out/r02_model_tests.log
for v in 0 1; do
LDX_FUSE_GN_STATS=$v python scripts/profile_unet.py --out gpurun_out/r02_ops_unet_gn$v.json > gpurun_out/r02_ops_unet_gn$v.txt 2>&1; tail -1 gpurun_out/r02_ops_unet_gn$v.txt
LDX_FUSE_GN_STATS=$v python scripts/profile_unet.py --vae --out gpurun_out/r02_ops_vae512_gn$v.json > gpurun_out/r02_ops_vae512_gn$v.txt 2>&1; tail -1 gpurun_out/r02_ops_vae512_gn$v.txt
done
LDX_FUSE_GN_STATS=0 python scripts/profile_unet.py --vae --batch 4 --latent 128 --out gpurun_out/r02_ops_vae1024_gn0.json > gpurun_out/r02_ops_vae1024_gn0.txt 2>&1; tail -1 gpurun_out/r02_ops_vae1024_gn0.txt
python scripts/profile_unet.py --vae --batch 4 --latent 128 --out gpurun_out/r02_ops_vae1024_gn1.json > gpurun_out/r02_ops_vae1024_gn1.txt 2>&1; tail -1 gpurun_out/r02_ops_vae1024_gn1.txt
echo "cross-attn NG=1"; python scripts/bench_cross_attn.py 2>&1 | tail -5
echo "cross-attn NG=4"; LDX_ATT_NG_SHORT=4 python scripts/bench_cross_attn.py 2>&1 | tail -5
