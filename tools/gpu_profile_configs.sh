#!/bin/bash
# ncu launch lists (duration + DRAM bytes) of the config-C3 and config-C5 steps at their BASELINE sizes
mkdir -p gpurun_out
for c in c3 c5; do
  timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_$c.csv python tools/profile_config.py $c 3 > gpurun_out/launches_$c.log 2>&1
  tail -1 gpurun_out/launches_$c.log
  python tools/launch_summary.py gpurun_out/launches_$c.csv 40 2>&1 | grep -v "at::" | head -32
done
