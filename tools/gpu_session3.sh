#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 -k "dgemm or fast_diag or stepper" > gpurun_out/pytest_gpu3.txt 2>&1
tail -5 gpurun_out/pytest_gpu3.txt
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_tma_4096x16384.json 2> gpurun_out/bench_tma.err; tail -c 1200 gpurun_out/bench_tma_4096x16384.json; tail -3 gpurun_out/bench_tma.err
NG=$(nvidia-smi -L | wc -l)
echo "gpus: $NG"
if [ "$NG" -ge 2 ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_slab.py 1024 6 > gpurun_out/check_slab2.txt 2>&1; tail -3 gpurun_out/check_slab2.txt
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -c 1200 gpurun_out/bench_2gpu.json; tail -3 gpurun_out/bench_2gpu.err
fi
