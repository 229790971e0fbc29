#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02aq}
timeout 600 python -m pytest tests/test_cuda_parity.py -x -q -m gpu -k "dct_rows or headline or fft_z" > gpurun_out/${TAG}_dct_tests.txt 2>&1
tail -3 gpurun_out/${TAG}_dct_tests.txt
timeout 120 python tools/bench_dct.py > gpurun_out/${TAG}_dct_bench.txt 2>&1
timeout 120 python tools/bench_dct.py 16384 512 >> gpurun_out/${TAG}_dct_bench.txt 2>&1
cat gpurun_out/${TAG}_dct_bench.txt
timeout 600 python bench.py --no-configs --no-cpu > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/${TAG}_bench_c4.json") if l.startswith("{")][-1])
print("c4", d["ms_per_step"], d["value"], d["roofline"]["frac"], d["roofline"]["solve_ms"], d["e2e"]["ms_per_step"])
PY
