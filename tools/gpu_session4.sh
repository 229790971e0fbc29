#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=900 --maxfail=30 > gpurun_out/pytest_gpu4.txt 2>&1
tail -40 gpurun_out/pytest_gpu4.txt | cut -c1-220
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_march_4096x16384.json 2> gpurun_out/bench_march.err; tail -c 900 gpurun_out/bench_march_4096x16384.json; tail -3 gpurun_out/bench_march.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r1c.csv python tools/profile_step.py 16384 2 > gpurun_out/launches_r1c.log 2>&1
tail -2 gpurun_out/launches_r1c.log
