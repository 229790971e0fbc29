#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=900 --maxfail=30 -k "parity_split or stepper" > gpurun_out/pytest_gpu11.txt 2>&1
tail -4 gpurun_out/pytest_gpu11.txt | cut -c1-220
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_slab.py 2048 6 > gpurun_out/check_slab2.txt 2>&1; tail -2 gpurun_out/check_slab2.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -c 900 gpurun_out/bench_2gpu.json; tail -2 gpurun_out/bench_2gpu.err
timeout 1500 ncu --set full --clock-control none --import-source on -k 'regex:k_dgemm|k_fd_fold' -s 16 -c 16 -o gpurun_out/prof_r1f python tools/profile_step.py 16384 2 > gpurun_out/prof_r1f.log 2>&1
tail -2 gpurun_out/prof_r1f.log
