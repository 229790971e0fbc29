#!/bin/bash
# ncu --set full of the round-2 kernels that are not HBM bound: lattice remesh gather, fused elastic pass, periodic FFT
mkdir -p gpurun_out
T=${TAG:-r02ai}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_p2m_lattice_gather|k_bubble_pre|k_cycle_avg3" -c 3 -o gpurun_out/${T}_ncu_c5 -f python tools/profile_config.py c5b 1 > gpurun_out/${T}_ncu_c5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_solid_fused|k_ls_fill|k_heav_mask" -c 3 -o gpurun_out/${T}_ncu_c3 -f python tools/profile_config.py c3d 1 > gpurun_out/${T}_ncu_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_rfft_rows|k_irfft_rows" -c 2 -o gpurun_out/${T}_ncu_c2 -f python tools/profile_config.py c2 1 > gpurun_out/${T}_ncu_c2.log 2>&1
for c in c5 c3 c2; do python tools/ncu_summary.py gpurun_out/${T}_ncu_$c.ncu-rep > gpurun_out/${T}_ncu_${c}_summary.csv; done
ls -la gpurun_out/${T}_ncu_*
