#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02o}
timeout 600 python -m pytest tests/test_cuda_parity.py -x -q -m gpu -k "dct_rows" > gpurun_out/${TAG}_dct_tests.txt 2>&1
tail -5 gpurun_out/${TAG}_dct_tests.txt
: > gpurun_out/${TAG}_sweep.txt
for v in ${SWEEP:-0 300 600 900 1200}; do
  echo "skew $v" >> gpurun_out/${TAG}_sweep.txt
  AXB_DCT_SKEW=$v timeout 120 python tools/bench_dct.py 2>&1 | grep dct >> gpurun_out/${TAG}_sweep.txt
done
timeout 120 python tools/bench_dct.py 16384 512 >> gpurun_out/${TAG}_sweep.txt
cat gpurun_out/${TAG}_sweep.txt
