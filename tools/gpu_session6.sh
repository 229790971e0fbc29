#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=900 --maxfail=30 > gpurun_out/pytest_gpu6.txt 2>&1
tail -8 gpurun_out/pytest_gpu6.txt | cut -c1-220
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r1d.csv python tools/profile_step.py 16384 2 > gpurun_out/launches_r1d.log 2>&1
tail -2 gpurun_out/launches_r1d.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r1d_4096.csv python tools/profile_step.py 4096 2 > gpurun_out/launches_r1d_4096.log 2>&1
