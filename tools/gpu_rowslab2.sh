#!/bin/bash
# 2-GPU validation of the r-slab stepper: tests, then the r-slab and z-slab bench lines
mkdir -p gpurun_out
TAG=${1:-r02i}
timeout 900 python -m pytest tests/test_rowslab_cuda.py tests/test_multigpu_cuda.py -x -q -m gpu > gpurun_out/${TAG}_rowslab_tests.txt 2>&1
tail -5 gpurun_out/${TAG}_rowslab_tests.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29555 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_2gpu_rows.json 2> gpurun_out/${TAG}_bench_2gpu_rows.err
timeout 300 $TR --master-port 29557 bench.py --gpus 2 --steps 10 --warmup 3 --no-graph > gpurun_out/${TAG}_bench_2gpu_rows_eager.json 2> gpurun_out/${TAG}_bench_2gpu_rows_eager.err
AXB_SLAB_Z=1 timeout 300 $TR --master-port 29556 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_2gpu_z.json 2> gpurun_out/${TAG}_bench_2gpu_z.err
AXB_GRAPH=1 timeout 200 $TR --master-port 29558 tools/rowslab_phases.py > gpurun_out/${TAG}_phases2.txt 2>&1
for f in rows rows_eager z; do python - <<P
import json
for l in open("gpurun_out/${TAG}_bench_2gpu_$f.json"):
    if l.startswith("{"):
        d=json.loads(l); print("$f", d["n_gpus"], round(d["ms_per_step"],4), d.get("slab_vs_single_rel_linf"), d.get("phases_ms"), d["e2e"]["ms_per_step"], d["gpu_launches"])
P
done
grep -h "^{" gpurun_out/${TAG}_phases2.txt | cut -c1-200
tail -3 gpurun_out/${TAG}_bench_2gpu_rows.err
