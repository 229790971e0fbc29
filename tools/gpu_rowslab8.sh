#!/bin/bash
# 8-GPU box: r-slab bench lines at N = 8, 4 (+ per-phase times)
mkdir -p gpurun_out
TAG=${1:-r02m}
for N in 8 4; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
  timeout 300 $TR --master-port 2955$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_${N}gpu_rows.json 2> gpurun_out/${TAG}_bench_${N}gpu_rows.err
  AXB_GRAPH=1 timeout 200 $TR --master-port 2956$N tools/rowslab_phases.py > gpurun_out/${TAG}_phases${N}.txt 2>&1
  python - <<P
import json
for l in open("gpurun_out/${TAG}_bench_${N}gpu_rows.json"):
    if l.startswith("{"):
        d=json.loads(l); print("rows", d["n_gpus"], round(d["ms_per_step"],4), d.get("slab_vs_single_rel_linf"), d.get("phases_ms"), d["e2e"]["ms_per_step"], d["roofline"]["solve_ms"])
P
  grep -h "^{" gpurun_out/${TAG}_phases${N}.txt | head -2 | cut -c1-160
done
tail -3 gpurun_out/${TAG}_bench_8gpu_rows.err
