#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=900 --maxfail=30 > gpurun_out/pytest_gpu13.txt 2>&1
tail -25 gpurun_out/pytest_gpu13.txt | cut -c1-200
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu --r-method tridiagonal > gpurun_out/bench_tri_4096x16384.json 2> gpurun_out/bench_tri.err; tail -c 1500 gpurun_out/bench_tri_4096x16384.json; tail -3 gpurun_out/bench_tri.err
timeout 600 python bench.py --config c1 --steps 50 --warmup 10 --no-cpu > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; tail -c 600 gpurun_out/bench_c1.json; tail -3 gpurun_out/bench_c1.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r1f.csv python tools/profile_step.py 16384 2 > gpurun_out/launches_r1f.log 2>&1
tail -1 gpurun_out/launches_r1f.log
