#!/bin/bash
# slim final check on one GPU (the last GPU-minutes of the round): full GPU suite, smoke, default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 --maxfail=20 > gpurun_out/r02_pytest_gpu.txt 2>&1
tail -2 gpurun_out/r02_pytest_gpu.txt | cut -c1-200
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r02_bench_1gpu.json") if l.startswith("{")][-1])
print("c4", d["ms_per_step"], d["value"], d["roofline"]["frac"], d["roofline"]["solve_ms"], d["step_roofline"]["frac"], (d.get("cpu_baseline") or {}).get("value"), d["e2e"]["ms_per_step"], d["gpu_launches"], d["clocks"])
for k, c in (d.get("configs") or {}).items():
    print(k, c.get("ms_per_step"), c.get("value"), (c.get("roofline") or {}).get("frac"), (c.get("step_roofline") or {}).get("frac"), (c.get("cpu_baseline") or {}).get("value"), (c.get("e2e") or {}).get("ms_per_step"))
PY
