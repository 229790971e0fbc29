#!/bin/bash
NG=8
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 tools/check_slab.py 2048 6 2>&1 | grep "slab x\|Error\|error" | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --steps 10 --warmup 3 2> gpurun_out/bench_${NG}gpu.err | grep '^{' > gpurun_out/bench_final_${NG}gpu.json; python -c "import json; d=json.load(open('gpurun_out/bench_final_${NG}gpu.json')); print('x$NG', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])"; tail -2 gpurun_out/bench_${NG}gpu.err
