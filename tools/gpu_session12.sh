#!/bin/bash
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l); echo "gpus: $NG"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 tools/check_slab.py 2048 6 > gpurun_out/check_slab$NG.txt 2>&1; tail -2 gpurun_out/check_slab$NG.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --steps 5 --warmup 3 2> gpurun_out/bench_${NG}gpu.err | grep '^{' > gpurun_out/bench_${NG}gpu.json; tail -c 700 gpurun_out/bench_${NG}gpu.json; tail -2 gpurun_out/bench_${NG}gpu.err
