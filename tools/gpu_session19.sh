#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=900 --maxfail=20 > gpurun_out/pytest_gpu19.txt 2>&1
tail -8 gpurun_out/pytest_gpu19.txt | cut -c1-220
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('fork', d['ms_per_step'], d['roofline']['solve_ms'], d['roofline']['frac'])"
AXB_EDGE_SERIAL=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('serial edges', d['ms_per_step'], d['roofline']['solve_ms'])"
timeout 300 python bench.py --config c1 --steps 50 --warmup 10 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c1', d['ms_per_step'], d['gpu_launches'])"
timeout 300 python bench.py --config c2 --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c2', d['ms_per_step'], d['roofline']['kernel'][:60])"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r1l.csv python tools/profile_step.py 16384 2 > gpurun_out/launches_r1l.log 2>&1
tail -1 gpurun_out/launches_r1l.log
