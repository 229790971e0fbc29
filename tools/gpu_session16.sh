#!/bin/bash
NG=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 --maxfail=20 -k "dct or fft or tridiag or fast_diag or stepper" > gpurun_out/pytest_gpu16.txt 2>&1
tail -15 gpurun_out/pytest_gpu16.txt | cut -c1-220
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_fft3_4096x16384.json 2> gpurun_out/bench_fft3.err; tail -c 1500 gpurun_out/bench_fft3_4096x16384.json; tail -3 gpurun_out/bench_fft3.err
for cols in 16 32 64; do AXB_TRI_COLS=$cols timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tri cols $cols', d['ms_per_step'], d['roofline']['solve_ms'])"; done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 tools/check_slab.py 2048 6 > gpurun_out/check_slab$NG.txt 2>&1; tail -3 gpurun_out/check_slab$NG.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --steps 10 --warmup 3 2> gpurun_out/bench_${NG}gpu.err | grep '^{' > gpurun_out/bench_${NG}gpu_fft.json; tail -c 900 gpurun_out/bench_${NG}gpu_fft.json; tail -2 gpurun_out/bench_${NG}gpu.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r1i.csv python tools/profile_step.py 16384 2 > gpurun_out/launches_r1i.log 2>&1
tail -1 gpurun_out/launches_r1i.log
