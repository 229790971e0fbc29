#!/bin/bash
# A/B of library builds that differ in a compile-time choice (here: __launch_bounds__ register caps of the interior
# row-marching kernels, profiles/r02_register_caps_ab.txt).  Build each variant here (edit, `make -C pyaxisymflow_b200/csrc`,
# copy libaxisym_b200.so to tools/_alt/<name>.so -- *.so is git-ignored but travels with gpurun), list the names below;
# the script swaps them in turn under the package and restores the original.  Last use: capreduce (penalisation kernel
# without the fused reduction left at 150 registers) against capall (capped at 128, 76 bytes spilled), C3 line three times each
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
