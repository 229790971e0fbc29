#!/bin/bash
# C3 / C5: side-stream overlaps (LS sweeps, psi averages) A/B; tests of the steppers
mkdir -p gpurun_out
T=${TAG:-r02bc}
timeout 600 python -m pytest tests -m gpu -q --timeout=300 --maxfail=10 -k "soft_sphere or particle or ensemble or restart or dct" > gpurun_out/${T}_pytest.txt 2>&1
tail -3 gpurun_out/${T}_pytest.txt | cut -c1-220
run() { # name, env...
  n=$1; shift
  env "$@" timeout 300 python bench.py --config ${n%%_*} --no-cpu > gpurun_out/${T}_bench_$n.json 2> gpurun_out/${T}_bench_$n.err
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/${T}_bench_$n.json') if l.startswith('{')][-1])
print('$n', d['ms_per_step'], d['value'], d['roofline'].get('solve_ms'), d.get('step_roofline',{}).get('frac'), d.get('gpu_launches'))"
}
run c3_overlap A=1
run c3_single AXB_SOFT_NO_OVERLAP=1
run c5 A=1
