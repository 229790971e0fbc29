#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_widen_cuda.py -q --timeout=200 2>&1 | tail -8 | cut -c1-300
timeout 200 python -m pytest tests/test_cuda_parity.py -q --timeout=180 -k "particle or soft_sphere" 2>&1 | tail -3
AXB_ENSEMBLE_SERIAL=1 timeout 200 python bench.py --config c5 --no-cpu --steps 10 --warmup 3 > gpurun_out/bench_c5_serial.json 2> gpurun_out/bench_c5_serial.err
timeout 200 python bench.py --config c5 --no-cpu --steps 10 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
python - <<'PY'
import json
for f in ("bench_c5_serial", "bench_c5"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json")); print(f, d["ms_per_step"], d["value"], d["config"]["workload"][-60:])
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/{f}.err").read()[-800:])
PY
