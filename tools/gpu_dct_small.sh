#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-r02al}
timeout 300 python -m pytest tests/test_cuda_parity.py tests/test_widen_cuda.py -q -m gpu -k "dct_rows or fft_z or heaviside" 2>&1 | tail -3
: > gpurun_out/${T}_dct_small.txt
for cfg in "8192 2048" "2048 8192" "4096 4096" "1024 4096"; do
  for v in 1 0; do
    if [ $v = 1 ]; then export AXB_DCT_2CTA=1; else unset AXB_DCT_2CTA; fi
    echo "N rows = $cfg  AXB_DCT_2CTA=${AXB_DCT_2CTA:-0}" >> gpurun_out/${T}_dct_small.txt
    timeout 120 python tools/bench_dct.py $cfg 2>&1 | grep dct >> gpurun_out/${T}_dct_small.txt
  done
done
cat gpurun_out/${T}_dct_small.txt
