#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-r02an}
timeout 900 python -m pytest tests/test_rowslab_cuda.py tests/test_multigpu_cuda.py -q -m gpu 2>&1 | tail -15 | cut -c1-400
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29561 bench.py --config c2 --gpus 2 --steps 20 --warmup 3 --no-cpu > gpurun_out/${T}_bench_c2_2gpu.json 2> gpurun_out/${T}_bench_c2_2gpu.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${T}_bench_c2_2gpu.json") if l.startswith("{")][-1])
    print("c2 x2", round(d["ms_per_step"], 4), d["value"], d.get("slab_vs_single_rel_linf"), d.get("slab_vs_single_case"), d.get("phases_ms"), d["config"]["parallelism"], d["gpu_launches"])
except Exception as e:
    print("c2 x2 failed", e); print(open("gpurun_out/${T}_bench_c2_2gpu.err").read()[-2500:])
PY
