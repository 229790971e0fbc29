#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=900 --maxfail=20 > gpurun_out/pytest_gpu25.txt 2>&1
tail -12 gpurun_out/pytest_gpu25.txt | cut -c1-220
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4', d['ms_per_step'], d['roofline']['solve_ms'], d['roofline']['frac'])"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r1p.csv python tools/profile_step.py 16384 2 > gpurun_out/launches_r1p.log 2>&1
