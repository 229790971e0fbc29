#!/bin/bash
NG=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 tools/check_slab.py 2048 6 > gpurun_out/check_slab${NG}_part.txt 2>&1; tail -3 gpurun_out/check_slab${NG}_part.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --steps 20 --warmup 5 2> gpurun_out/bench_${NG}gpu.err | grep '^{' > gpurun_out/bench_${NG}gpu_part.json; python -c "import json; d=json.load(open('gpurun_out/bench_${NG}gpu_part.json')); print('partition x$NG', d['ms_per_step'], d['roofline']['solve_ms'], d['value'])"; tail -2 gpurun_out/bench_${NG}gpu.err
AXB_SLAB_TRANSPOSE4=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $NG --steps 20 --warmup 5 2> gpurun_out/bench_${NG}gpu_t4.err | grep '^{' > gpurun_out/bench_${NG}gpu_t4.json; python -c "import json; d=json.load(open('gpurun_out/bench_${NG}gpu_t4.json')); print('4 transposes x$NG', d['ms_per_step'], d['roofline']['solve_ms'])"
