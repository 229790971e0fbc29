#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=900 --maxfail=20 > gpurun_out/pytest_gpu23.txt 2>&1
tail -4 gpurun_out/pytest_gpu23.txt | cut -c1-220
timeout 900 python bench.py > gpurun_out/bench_r01_final_c4.json 2> gpurun_out/bench_r01_final_c4.err; tail -c 2500 gpurun_out/bench_r01_final_c4.json; tail -3 gpurun_out/bench_r01_final_c4.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_reference_c4.json 2> gpurun_out/bench_r01_reference.err; tail -c 900 gpurun_out/bench_r01_reference_c4.json; tail -3 gpurun_out/bench_r01_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r1n.csv python tools/profile_step.py 16384 2 > gpurun_out/launches_r1n.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:'k_dct_rows_rr|k_tri_sweep_tma|km_penalise' --launch-skip 6 --launch-count 6 -o gpurun_out/r01_solve_full -f python tools/profile_step.py 16384 2 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
