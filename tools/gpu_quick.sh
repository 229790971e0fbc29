#!/bin/bash
# quick check: selected GPU tests (pytest -k "$1", default: the solve), C4 bench, ncu launch list
mkdir -p gpurun_out
K=${1:-"dct or fft or fast_diag"}
timeout 900 python -m pytest tests -m gpu -q --timeout=600 --maxfail=20 -k "$K" > gpurun_out/pytest_quick.txt 2>&1
tail -3 gpurun_out/pytest_quick.txt | cut -c1-220
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4', d['ms_per_step'], d['roofline']['solve_ms'], d['roofline']['frac'])"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_quick.csv python tools/profile_step.py 16384 2 > gpurun_out/launches_quick.log 2>&1
