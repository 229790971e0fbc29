#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-r02as}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"km_penalise|km_eno3|km_diffusion_fused|km_velocity" -c 8 -o gpurun_out/${T}_ncu_stencils -f python tools/profile_step.py 16384 1 > gpurun_out/${T}_ncu_stencils.log 2>&1
python tools/ncu_summary.py gpurun_out/${T}_ncu_stencils.ncu-rep > gpurun_out/${T}_ncu_stencils_summary.csv
ls -la gpurun_out/${T}_*
