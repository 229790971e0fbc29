#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 --maxfail=20 -k "dct or fft or tridiag or fast_diag" > gpurun_out/pytest_gpu17.txt 2>&1
tail -5 gpurun_out/pytest_gpu17.txt | cut -c1-220
for cols in 16 32 64 128; do AXB_TRI_COLS=$cols timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tri cols $cols', d['ms_per_step'], d['roofline']['solve_ms'])"; done
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r1j.csv python tools/profile_step.py 16384 2 > gpurun_out/launches_r1j.log 2>&1
tail -1 gpurun_out/launches_r1j.log
