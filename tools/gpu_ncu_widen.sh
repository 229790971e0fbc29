#!/bin/bash
# ncu --set full of the 8f kernels: one re-initialisation at 2048x8192 and the widen micro-benchmark kernels
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_reinit_sweep|k_reinit_front|k_reinit_ring" -c 8 -o gpurun_out/ncu_reinit -f python tools/profile_reinit.py 1 > gpurun_out/ncu_reinit.log 2>&1
tail -2 gpurun_out/ncu_reinit.log
ncu -i gpurun_out/ncu_reinit.ncu-rep --page raw --csv > gpurun_out/ncu_reinit_raw.csv 2>/dev/null
ls -la gpurun_out/ncu_reinit*
