#!/bin/bash
# multi-GPU bench lines (run under gpurun --gpus N): weak scaling (8 images per GPU) and BASELINE config 3 (64 images total)
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02c_bench_n$N.json 2> gpurun_out/r02c_bench_n$N.err; echo "weak rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --total-images 64 > gpurun_out/r02c_bench_n${N}_total64.json 2> gpurun_out/r02c_bench_n${N}_total64.err; echo "total64 rc=$?"
python - $N <<'P'
import json,sys
n=sys.argv[1]
for f in (f'gpurun_out/r02c_bench_n{n}.json', f'gpurun_out/r02c_bench_n{n}_total64.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'ms', round(d['ms_per_step'],1), d['scaling'], d['config']['images_per_gpu'], d['e2e']['d2h_bytes_per_step'])
    except Exception as e:
        print(f, 'FAILED', e)
P
tail -5 gpurun_out/r02c_bench_n$N.err
