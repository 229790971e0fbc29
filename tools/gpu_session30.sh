#!/bin/bash
NG=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --steps 20 --warmup 5 2> gpurun_out/bench_${NG}gpu.err | grep '^{' > gpurun_out/bench_${NG}gpu_e2e.json; python -c "import json; d=json.load(open('gpurun_out/bench_${NG}gpu_e2e.json')); print('x$NG', d['ms_per_step'], d['value'], d['e2e'])"; tail -2 gpurun_out/bench_${NG}gpu.err
