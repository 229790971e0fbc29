#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-r02ak}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_dct2_rows|k_dct3_rows|k_tri_sweep_tma" -c 4 -o gpurun_out/${T}_ncu_dct_c3 -f python tools/profile_config.py c3d 1 > gpurun_out/${T}_ncu_dct_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_dct2_rows|k_dct3_rows" -c 2 -o gpurun_out/${T}_ncu_dct_c5 -f python tools/profile_config.py c5b 1 > gpurun_out/${T}_ncu_dct_c5.log 2>&1
for c in c3 c5; do python tools/ncu_summary.py gpurun_out/${T}_ncu_dct_$c.ncu-rep > gpurun_out/${T}_ncu_dct_${c}_summary.csv; done
ncu -i gpurun_out/${T}_ncu_dct_c3.ncu-rep --page source --csv -k regex:k_dct2_rows 2>/dev/null | head -400 > gpurun_out/${T}_dct2_source.csv
ls -la gpurun_out/${T}_*
