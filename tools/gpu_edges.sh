#!/bin/bash
# edge kernels of the row-marching stencil passes: compact edge grid + fine row chunks (AXB_EDGE_RB) against the old
# full grid with RB-row edge blocks (AXB_EDGE_FULL_GRID=1 AXB_EDGE_RB=32)
mkdir -p gpurun_out
T=${TAG:-r02bg}
timeout 900 python -m pytest tests/test_cuda_parity.py tests/test_rowslab_cuda.py tests/test_widen_cuda.py -m gpu -q -x --timeout=600 > gpurun_out/${T}_pytest.txt 2>&1
tail -3 gpurun_out/${T}_pytest.txt | cut -c1-220
line() { python -c "
import json,sys
d=json.loads([l for l in open('$1') if l.startswith('{')][-1])
print('$2', round(d['ms_per_step'],4), d['value'])"; }
for v in "old AXB_EDGE_FULL_GRID=1 AXB_EDGE_RB=32" "c32 AXB_EDGE_RB=32" "c8 AXB_EDGE_RB=8" "c4 AXB_EDGE_RB=4"; do
  set -- $v; name=$1; shift
  for c in c4 c2 c3 c5 c1; do
    env "$@" timeout 300 python bench.py --config $c --no-cpu --no-configs > gpurun_out/${T}_bench_${c}_$name.json 2> gpurun_out/${T}_bench_${c}_$name.err
    line gpurun_out/${T}_bench_${c}_$name.json "$c $name"
  done
  env "$@" timeout 300 python bench.py --nr 516 --nz 16384 --no-cpu --no-configs > gpurun_out/${T}_bench_slab516_$name.json 2> gpurun_out/${T}_bench_slab516_$name.err
  line gpurun_out/${T}_bench_slab516_$name.json "516x16384 $name"
done
