#!/bin/bash
NG=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 tools/check_slab.py 2048 6 > gpurun_out/check_slab$NG.txt 2>&1; tail -4 gpurun_out/check_slab$NG.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --steps 10 --warmup 3 2> gpurun_out/bench_${NG}gpu.err | grep '^{' > gpurun_out/bench_${NG}gpu_p2p.json; python -c "import json; d=json.load(open('gpurun_out/bench_${NG}gpu_p2p.json')); print('p2p x$NG', d['ms_per_step'], d['roofline']['solve_ms'], d['roofline']['kernel'][-90:])"; tail -2 gpurun_out/bench_${NG}gpu.err
AXB_SLAB_NCCL=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $NG --steps 10 --warmup 3 2> gpurun_out/bench_${NG}gpu_nccl.err | grep '^{' > gpurun_out/bench_${NG}gpu_nccl.json; python -c "import json; d=json.load(open('gpurun_out/bench_${NG}gpu_nccl.json')); print('nccl x$NG', d['ms_per_step'], d['roofline']['solve_ms'])"
