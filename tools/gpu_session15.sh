#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 --maxfail=20 -k "dct or fft or tridiag or fast_diag or stepper" > gpurun_out/pytest_gpu15.txt 2>&1
tail -25 gpurun_out/pytest_gpu15.txt | cut -c1-220
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_fft2_4096x16384.json 2> gpurun_out/bench_fft2.err; tail -c 1800 gpurun_out/bench_fft2_4096x16384.json; tail -3 gpurun_out/bench_fft2.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r1h.csv python tools/profile_step.py 16384 2 > gpurun_out/launches_r1h.log 2>&1
tail -1 gpurun_out/launches_r1h.log
