#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-r02af}
timeout 600 python -m pytest tests/test_widen_cuda.py tests/test_cuda_parity.py -q --timeout=300 -m gpu -k "restart or host_step_pipeline or ensemble or device_scalars" 2>&1 | tail -8 | cut -c1-300
timeout 600 python bench.py --no-configs --no-cpu > gpurun_out/${T}_bench_c4.json 2> gpurun_out/${T}_bench_c4.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/${T}_bench_c4.json") if l.startswith("{")][-1])
print("c4", d["ms_per_step"], d["roofline"]["frac"], json.dumps(d["e2e"])[:700])
PY
tail -2 gpurun_out/${T}_bench_c4.err
