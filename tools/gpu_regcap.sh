#!/bin/bash
# A/B of register caps (__launch_bounds__(128, N)) on the interior row-marching kernels: prebuilt libraries under
# tools/_alt (both = penalise + fused RK2 at 128 registers; all4 = + ENO3 and velocity at 80), C4 twice, C3 and C5 once
mkdir -p gpurun_out
T=${TAG:-r02bk}
cp pyaxisymflow_b200/libaxisym_b200.so /tmp/orig.so
line() { python -c "
import json,sys
d=json.loads([l for l in open('$1') if l.startswith('{')][-1])
print('$2', round(d['ms_per_step'],4), d['value'])"; }
for v in both all4 both all4; do
  cp tools/_alt/$v.so pyaxisymflow_b200/libaxisym_b200.so
  timeout 300 python bench.py --no-cpu --no-configs --steps 20 --warmup 5 > gpurun_out/${T}_c4_$v.json 2> gpurun_out/${T}_c4_$v.err
  line gpurun_out/${T}_c4_$v.json "c4 $v"
done
for c in c3 c5; do
for v in both all4; do
  cp tools/_alt/$v.so pyaxisymflow_b200/libaxisym_b200.so
  timeout 300 python bench.py --config $c --no-cpu --no-configs > gpurun_out/${T}_${c}_$v.json 2> gpurun_out/${T}_${c}_$v.err
  line gpurun_out/${T}_${c}_$v.json "$c $v"
done
done
cp tools/_alt/all4.so pyaxisymflow_b200/libaxisym_b200.so
timeout 300 python -m pytest tests/test_cuda_parity.py -m gpu -q -x -k "penalis or diffusion or stepper or eno3 or velocity" --timeout=300 2>&1 | tail -2
cp /tmp/orig.so pyaxisymflow_b200/libaxisym_b200.so
