#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02q}
: > gpurun_out/${TAG}_sweep.txt
for v in ${SWEEP:-0 200 400 600 800 1000 1300 1600 2000}; do
  echo "skew $v" >> gpurun_out/${TAG}_sweep.txt
  AXB_DCT_SKEW=$v timeout 120 python tools/bench_dct.py 2>&1 | grep dct2 >> gpurun_out/${TAG}_sweep.txt
done
cat gpurun_out/${TAG}_sweep.txt
