#!/bin/bash
# config C5: batched / device-resident ensemble tests, then the c5 line in its three modes
mkdir -p gpurun_out
T=${TAG:-r02w}
timeout 600 python -m pytest tests/test_widen_cuda.py -q --timeout=300 -m gpu 2>&1 | tail -15 | cut -c1-400 > gpurun_out/${T}_widen_tests.txt
timeout 200 python -m pytest tests/test_cuda_parity.py -q --timeout=180 -m gpu -k "particle or soft_sphere or p2m or remesh" 2>&1 | tail -15 | cut -c1-300 >> gpurun_out/${T}_widen_tests.txt
for mode in streams members batched; do
  AXB_ENSEMBLE=$mode timeout 300 python bench.py --config c5 --no-cpu --steps 20 --warmup 3 > gpurun_out/${T}_bench_c5_$mode.json 2> gpurun_out/${T}_bench_c5_$mode.err
done
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_c5b.csv python tools/profile_config.py c5b 3 > gpurun_out/${T}_launches_c5b.log 2>&1
python tools/launch_summary.py gpurun_out/${T}_launches_c5b.csv 40 2>&1 | grep -v "at::" | head -32 > gpurun_out/${T}_kernels_c5.txt
cat gpurun_out/${T}_widen_tests.txt
python - <<PY
import json
for f in ("streams", "members", "batched"):
    p = f"gpurun_out/${T}_bench_c5_{f}"
    try:
        d = json.loads([l for l in open(p + ".json") if l.startswith("{")][-1]); print(f, d["ms_per_step"], d["value"], d.get("e2e", {}).get("ms_per_step"))
    except Exception as e:
        print(f, "failed", e); print(open(p + ".err").read()[-1500:])
PY
tail -40 gpurun_out/${T}_kernels_c5.txt
