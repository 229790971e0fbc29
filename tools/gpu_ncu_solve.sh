#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-r02ar}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_dct_rows_w" -c 2 -o gpurun_out/${T}_ncu_dct_w -f python tools/profile_step.py 16384 1 > gpurun_out/${T}_ncu_dct_w.log 2>&1
python tools/ncu_summary.py gpurun_out/${T}_ncu_dct_w.ncu-rep > gpurun_out/${T}_ncu_dct_w_summary.csv
cp pyaxisymflow_b200/libaxisym_b200.so gpurun_out/${T}_libaxisym_b200.so.copy 2>/dev/null
ls -la gpurun_out/${T}_*
