#!/bin/bash
# first GPU session: FP64 pipe probe, parity tests, smoke, small + full bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu_info.txt 2>&1
nproc >> gpurun_out/gpu_info.txt; lscpu | grep "Model name" >> gpurun_out/gpu_info.txt
timeout 120 ./tools/probe_dmma > gpurun_out/probe_dmma.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --maxfail=60 -x --timeout=600 > gpurun_out/pytest_gpu.txt 2>&1
tail -5 gpurun_out/pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.txt 2>&1; tail -2 gpurun_out/smoke.txt
timeout 600 python bench.py --nz 4096 --steps 5 --warmup 3 > gpurun_out/bench_1024x4096.json 2> gpurun_out/bench_1024x4096.err; tail -c 1500 gpurun_out/bench_1024x4096.json; tail -3 gpurun_out/bench_1024x4096.err
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_4096x16384.json 2> gpurun_out/bench_4096x16384.err; tail -c 2500 gpurun_out/bench_4096x16384.json; tail -3 gpurun_out/bench_4096x16384.err
