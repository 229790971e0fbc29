#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_reinit.csv python tools/profile_reinit.py 2 > gpurun_out/launches_reinit.log 2>&1
tail -2 gpurun_out/launches_reinit.log
python tools/launch_summary.py gpurun_out/launches_reinit.csv 2>&1 | tail -20
