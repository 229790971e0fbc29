#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-r02at}
timeout 600 python -m pytest tests/test_widen_cuda.py tests/test_cuda_parity.py -q --timeout=300 -m gpu -k "remesh or p2m or ensemble or particle" 2>&1 | tail -4 | cut -c1-300
timeout 300 python bench.py --config c5 --no-cpu --steps 20 --warmup 3 > gpurun_out/${T}_bench_c5.json 2> gpurun_out/${T}_bench_c5.err
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_c5b.csv python tools/profile_config.py c5b 3 > gpurun_out/${T}_launches_c5b.log 2>&1
python tools/launch_summary.py gpurun_out/${T}_launches_c5b.csv 50 2>&1 | grep -v "at::" > gpurun_out/${T}_kernel_summary_c5b.txt
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/${T}_bench_c5.json") if l.startswith("{")][-1])
print("c5", d["ms_per_step"], d["value"], d["roofline"]["frac"], d.get("step_roofline", {}).get("frac"))
PY
head -6 gpurun_out/${T}_kernel_summary_c5b.txt
