#!/bin/bash
# round 2, third session, final state: whole suite, smoke, both bench arms, ncu launch list of one UNet step and a
# --set full capture of the fused feed-forward kernel (each ncu command follows a plain run of the same command line)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02c_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -2 gpurun_out/r02c_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02c_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02c_smoke.log
python bench.py > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02c_bench_reference_arm.json 2> gpurun_out/r02c_bench_ref.err; echo "ref rc=$?"
python scripts/profile_unet.py --out gpurun_out/r02c_ops_unet.json > gpurun_out/r02c_ops_unet.txt 2>&1; tail -1 gpurun_out/r02c_ops_unet.txt
F='regex:ldx_|gn_|splitk_|cfg_|nchw_|timestep_|eps_to'
python scripts/profile_unet.py --eager-only > gpurun_out/ncu_plain_unet.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$F" -c 1600 --csv \
    --log-file gpurun_out/r02c_unet_step_launches.csv python scripts/profile_unet.py --eager-only > gpurun_out/ncu_unet.log 2>&1
echo "launch list rc=$?"
python scripts/ncu_targets.py ff_fused > gpurun_out/ncu_plain_ff_fused.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:ldx_ff_fused_kernel' -s 2 -c 2 -f -o gpurun_out/r02c_ncu_ff_fused python scripts/ncu_targets.py ff_fused > gpurun_out/ncu_ff_fused.log 2>&1
echo "ff_fused ncu rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02c_bench_n1.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'unet', d['unet_step_ms'], 'roofline', d['roofline']['frac'], d['roofline']['frac_executed'], d['roofline']['kernels'], 'gpu_ref', d.get('gpu_reference'), d['unet_breakdown'], d['clocks'])
print(open('gpurun_out/r02c_bench_reference_arm.json').read()[-600:])
P
