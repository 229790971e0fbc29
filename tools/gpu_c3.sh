#!/bin/bash
# config C3: device-resident soft-sphere step -- tests, the c3 line host-driven / device-resident, ncu launch list
mkdir -p gpurun_out
T=${TAG:-r02z}
timeout 600 python -m pytest tests/test_widen_cuda.py tests/test_cuda_parity.py -q --timeout=300 -m gpu -k "soft_sphere or ls_extrap or least_squares or cycle_averages or restart or solid" 2>&1 | tail -15 | cut -c1-400 > gpurun_out/${T}_c3_tests.txt
AXB_SOFT_HOST=1 timeout 300 python bench.py --config c3 --no-cpu --steps 10 --warmup 3 > gpurun_out/${T}_bench_c3_host.json 2> gpurun_out/${T}_bench_c3_host.err
timeout 300 python bench.py --config c3 --no-cpu --steps 10 --warmup 3 > gpurun_out/${T}_bench_c3_dev.json 2> gpurun_out/${T}_bench_c3_dev.err
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_c3d.csv python tools/profile_config.py c3d 3 > gpurun_out/${T}_launches_c3d.log 2>&1
python tools/launch_summary.py gpurun_out/${T}_launches_c3d.csv 50 2>&1 | grep -v "at::" | head -45 > gpurun_out/${T}_kernels_c3.txt
cat gpurun_out/${T}_c3_tests.txt
python - <<PY
import json
for f in ("host", "dev"):
    p = f"gpurun_out/${T}_bench_c3_{f}"
    try:
        d = json.loads([l for l in open(p + ".json") if l.startswith("{")][-1]); print(f, d["ms_per_step"], d["value"], d.get("e2e", {}).get("ms_per_step"))
    except Exception as e:
        print(f, "failed", e); print(open(p + ".err").read()[-1500:])
PY
head -24 gpurun_out/${T}_kernels_c3.txt
