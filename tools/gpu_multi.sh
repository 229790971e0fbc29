#!/bin/bash
# multi-GPU validation on an N-GPU box: tools/gpu_multi.sh TAG N1 [N2 ...]
#   r-slab tests, then per N: the default bench line (r-slabs, one graph per rank, slab-vs-single check, phases),
#   the reference arm under torchrun, and -- for the largest N -- the C5 ensemble as N independent batched replicas
mkdir -p gpurun_out
TAG=${1:-r02}; shift
timeout 900 python -m pytest tests/test_rowslab_cuda.py tests/test_multigpu_cuda.py -q -m gpu > gpurun_out/${TAG}_multigpu_tests.txt 2>&1
tail -3 gpurun_out/${TAG}_multigpu_tests.txt
LAST=1
for N in "$@"; do
  LAST=$N
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
  timeout 400 $TR --master-port 2955$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
  python - <<P
import json
try:
    d = json.loads([l for l in open("gpurun_out/${TAG}_bench_${N}gpu.json") if l.startswith("{")][-1])
    print("c4 x$N", round(d["ms_per_step"], 4), d["value"], d.get("slab_vs_single_rel_linf"), d.get("phases_ms"), d["e2e"]["ms_per_step"], d["roofline"]["solve_ms"], d["gpu_launches"], d["clocks"])
except Exception as e:
    print("c4 x$N failed", e); print(open("gpurun_out/${TAG}_bench_${N}gpu.err").read()[-1500:])
P
done
N=$LAST
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29571 bench.py --config c5 --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_c5_${N}gpu.json 2> gpurun_out/${TAG}_bench_c5_${N}gpu.err
timeout 400 $TR --master-port 29572 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm_${N}gpu.json 2> /dev/null
python - <<P
import json
try:
    d = json.loads([l for l in open("gpurun_out/${TAG}_bench_c5_${N}gpu.json") if l.startswith("{")][-1])
    print("c5 x$N", round(d["ms_per_step"], 4), d["value"], d["config"], d["gpu_launches"])
except Exception as e:
    print("c5 x$N failed", e); print(open("gpurun_out/${TAG}_bench_c5_${N}gpu.err").read()[-1500:])
P
cut -c1-200 gpurun_out/${TAG}_bench_reference_arm_${N}gpu.json
