#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=900 --maxfail=20 > gpurun_out/pytest_gpu26.txt 2>&1
tail -4 gpurun_out/pytest_gpu26.txt | cut -c1-220
timeout 900 python bench.py > gpurun_out/bench_r01i_c4.json 2> gpurun_out/bench_r01i_c4.err; tail -c 2600 gpurun_out/bench_r01i_c4.json; tail -3 gpurun_out/bench_r01i_c4.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r1q.csv python tools/profile_step.py 16384 2 > gpurun_out/launches_r1q.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:'k_dct_rows_rr|k_tri_sweep_tma|km_diffusion_fused' --launch-skip 5 --launch-count 5 -o gpurun_out/r01i_solve_full -f python tools/profile_step.py 16384 2 > gpurun_out/ncu_full2.log 2>&1
tail -1 gpurun_out/ncu_full2.log
