#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-r02ag}
timeout 600 python -m pytest tests/test_cuda_parity.py -q --timeout=300 -m gpu -k "periodic or rfft or fast_diagonalisation" 2>&1 | tail -6 | cut -c1-300
timeout 300 python bench.py --config c2 --no-cpu --steps 20 --warmup 3 > gpurun_out/${T}_bench_c2.json 2> gpurun_out/${T}_bench_c2.err
timeout 300 python bench.py --config c1 --no-cpu --steps 20 --warmup 3 > gpurun_out/${T}_bench_c1.json 2> gpurun_out/${T}_bench_c1.err
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_c2.csv python tools/profile_config.py c2 3 > gpurun_out/${T}_launches_c2.log 2>&1
python tools/launch_summary.py gpurun_out/${T}_launches_c2.csv 30 2>&1 | grep -v "at::" | head -14 > gpurun_out/${T}_kernel_summary_c2.txt
python - <<PY
import json
for c in ("c2", "c1"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/${T}_bench_{c}.json") if l.startswith("{")][-1])
        print(c, d["ms_per_step"], d["value"], d["roofline"]["frac"], d["roofline"]["solve_ms"], d["e2e"]["ms_per_step"], d.get("step_roofline", {}).get("frac"), d["gpu_launches"])
    except Exception as e:
        print(c, "failed", e); print(open(f"gpurun_out/${T}_bench_{c}.err").read()[-1500:])
PY
cat gpurun_out/${T}_kernel_summary_c2.txt
