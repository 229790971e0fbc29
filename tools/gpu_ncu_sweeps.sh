#!/bin/bash
# ncu --set full of the warp-specialised sweeps at the C2 shape (1024 x 4096: 128 CTAs), one and three producer warps
mkdir -p gpurun_out
T=${TAG:-r02ba}
AXB_TRI_RING=8 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_tri_sweep_ws" -c 2 -o gpurun_out/${T}_ncu_sweeps_c2_np1 -f python tools/profile_sweeps.py 1024 4096 1 > gpurun_out/${T}_ncu_sweeps_np1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_tri_sweep_ws" -c 2 -o gpurun_out/${T}_ncu_sweeps_c2_np3 -f python tools/profile_sweeps.py 1024 4096 1 > gpurun_out/${T}_ncu_sweeps_np3.log 2>&1
ls -la gpurun_out/${T}_*
