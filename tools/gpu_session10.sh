#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=900 --maxfail=30 > gpurun_out/pytest_gpu10.txt 2>&1
tail -30 gpurun_out/pytest_gpu10.txt | cut -c1-220
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_split_4096x16384.json 2> gpurun_out/bench_split.err; tail -c 2600 gpurun_out/bench_split_4096x16384.json; tail -3 gpurun_out/bench_split.err
