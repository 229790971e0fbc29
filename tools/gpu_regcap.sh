#!/bin/bash
# A/B of the register cap on the interior penalisation kernel WITHOUT the fused reduction (config C3's form): prebuilt
# libraries under tools/_alt (capreduce = 150 registers, capall = 128 with 76 bytes spilled), C3 line three times each
mkdir -p gpurun_out
T=${TAG:-r02bn}
cp pyaxisymflow_b200/libaxisym_b200.so /tmp/orig.so
for v in capreduce capall capreduce capall capreduce capall; do
  cp tools/_alt/$v.so pyaxisymflow_b200/libaxisym_b200.so
  timeout 200 python bench.py --config c3 --no-cpu --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c3 $v', round(d['ms_per_step'],4), d['value'])"
done
cp /tmp/orig.so pyaxisymflow_b200/libaxisym_b200.so
