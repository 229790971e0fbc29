#!/bin/bash
# ncu launch list (+ DRAM bytes) of three C4 steps and its per-kernel aggregate
mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_launches_c4.csv python tools/profile_step.py 16384 3 > gpurun_out/r02_launches_c4.log 2>&1
python tools/launch_summary.py gpurun_out/r02_launches_c4.csv 40 2>&1 | grep -v "at::" | head -40 > gpurun_out/r02_kernel_summary_c4.txt
cat gpurun_out/r02_kernel_summary_c4.txt | head -16
