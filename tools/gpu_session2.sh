#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.txt 2>&1; tail -2 gpurun_out/smoke.txt
# launch list (cold-cache, serialised): shares only
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python tools/profile_step.py 16384 2 > gpurun_out/launches_r1.log 2>&1
tail -2 gpurun_out/launches_r1.log
# full capture of the second step's heavy kernels
timeout 1500 ncu --set full --clock-control none --import-source on -k 'regex:k_dgemm|k_eno3|k_penalise|k_velocity|k_diffusion' -s 9 -c 9 -o gpurun_out/prof_r1 python tools/profile_step.py 16384 2 > gpurun_out/prof_r1.log 2>&1
tail -3 gpurun_out/prof_r1.log; ls -la gpurun_out/
