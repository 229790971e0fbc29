#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=900 --maxfail=30 > gpurun_out/pytest_gpu9.txt 2>&1
tail -5 gpurun_out/pytest_gpu9.txt | cut -c1-220
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r1e.csv python tools/profile_step.py 16384 2 > gpurun_out/launches_r1e.log 2>&1
tail -1 gpurun_out/launches_r1e.log
for cfg in c2 c3 c5; do
  timeout 900 python bench.py --config $cfg --steps 5 --warmup 3 > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err
  tail -2 gpurun_out/bench_$cfg.err
done
