#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 --maxfail=20 -k "dct or fft or tridiag or fast_diag or stepper" > gpurun_out/pytest_gpu20.txt 2>&1
tail -8 gpurun_out/pytest_gpu20.txt | cut -c1-220
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('rr', d['ms_per_step'], d['roofline']['solve_ms'], d['roofline']['frac'])"
AXB_DCT_SMEM=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('smem dct', d['ms_per_step'], d['roofline']['solve_ms'])"
timeout 300 python bench.py --config c1 --steps 50 --warmup 10 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c1', d['ms_per_step'], d['gpu_launches'])"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r1m.csv python tools/profile_step.py 16384 2 > gpurun_out/launches_r1m.log 2>&1
tail -1 gpurun_out/launches_r1m.log
