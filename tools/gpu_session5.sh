#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=900 --maxfail=30 > gpurun_out/pytest_gpu5.txt 2>&1
tail -8 gpurun_out/pytest_gpu5.txt | cut -c1-220
timeout 1500 ncu --set full --clock-control none --import-source on -k 'regex:km_' -s 5 -c 5 -o gpurun_out/prof_r1c python tools/profile_step.py 16384 2 > gpurun_out/prof_r1c.log 2>&1
tail -2 gpurun_out/prof_r1c.log
