#!/bin/bash
# warp-specialised tridiagonal sweeps + occupancy-sized DCT grids: parity tests, micro-benchmarks, C2 / C3 / C5 lines
mkdir -p gpurun_out
T=${TAG:-r02az}
timeout 600 python -m pytest tests -m gpu -q --timeout=300 --maxfail=10 -k "tridiagonal or fast_diag or headline or rowslab or potential or dct" > gpurun_out/${T}_pytest_sweeps.txt 2>&1
tail -3 gpurun_out/${T}_pytest_sweeps.txt | cut -c1-220
timeout 300 python tools/bench_sweeps.py > gpurun_out/${T}_bench_sweeps.txt 2>&1
cat gpurun_out/${T}_bench_sweeps.txt | tail -8 | cut -c1-150
for r in 8 16 32; do echo "ring $r"; AXB_TRI_RING=$r timeout 300 python tools/bench_sweeps.py 2>&1 | cut -c1-120 | tail -6; done > gpurun_out/${T}_bench_sweeps_rings.txt 2>&1
cat gpurun_out/${T}_bench_sweeps_rings.txt
for sz in "8192 2048" "2048 8192" "4096 4096"; do
  echo "occupancy grid: $(timeout 120 python tools/bench_dct.py $sz 2>&1 | tr '\n' ' ')"
  echo "smem grid:      $(AXB_DCT_GRID_SMEM=1 timeout 120 python tools/bench_dct.py $sz 2>&1 | tr '\n' ' ')"
done > gpurun_out/${T}_dct_grid_ab.txt 2>&1
cat gpurun_out/${T}_dct_grid_ab.txt
for c in c2 c3 c5; do
  timeout 300 python bench.py --config $c --no-cpu > gpurun_out/${T}_bench_${c}.json 2> gpurun_out/${T}_bench_${c}.err
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/${T}_bench_${c}.json') if l.startswith('{')][-1])
print('$c', d['ms_per_step'], d['value'], d['roofline'].get('solve_ms'), d.get('step_roofline',{}).get('frac'))"
done
timeout 300 python bench.py --no-cpu --no-configs --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4', d['ms_per_step'], d['roofline']['solve_ms'], d['roofline']['frac'])"
