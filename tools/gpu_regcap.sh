#!/bin/bash
# A/B of register caps (__launch_bounds__(128, 4)) on km_penalise<.,1> and km_diffusion_fused<1>: four prebuilt libraries
# under tools/_alt (nocap / pen / rk2 / both), C4 line twice and C5 once per variant
mkdir -p gpurun_out
T=${TAG:-r02bj}
cp pyaxisymflow_b200/libaxisym_b200.so /tmp/orig.so
line() { python -c "
import json,sys
d=json.loads([l for l in open('$1') if l.startswith('{')][-1])
print('$2', round(d['ms_per_step'],4), d['value'])"; }
for v in nocap pen rk2 both nocap both; do
  cp tools/_alt/$v.so pyaxisymflow_b200/libaxisym_b200.so
  timeout 300 python bench.py --no-cpu --no-configs --steps 20 --warmup 5 > gpurun_out/${T}_c4_$v.json 2> gpurun_out/${T}_c4_$v.err
  line gpurun_out/${T}_c4_$v.json "c4 $v"
done
for v in nocap pen; do
  cp tools/_alt/$v.so pyaxisymflow_b200/libaxisym_b200.so
  timeout 300 python bench.py --config c5 --no-cpu --no-configs > gpurun_out/${T}_c5_$v.json 2> gpurun_out/${T}_c5_$v.err
  line gpurun_out/${T}_c5_$v.json "c5 $v"
done
cp tools/_alt/both.so pyaxisymflow_b200/libaxisym_b200.so
timeout 300 python -m pytest tests/test_cuda_parity.py -m gpu -q -x -k "penalis or diffusion or stepper" --timeout=300 2>&1 | tail -2
cp /tmp/orig.so pyaxisymflow_b200/libaxisym_b200.so
