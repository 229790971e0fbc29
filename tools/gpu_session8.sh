#!/bin/bash
mkdir -p gpurun_out
for cfg in c2 c3 c5; do
  timeout 900 python bench.py --config $cfg --steps 5 --warmup 3 > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err
  tail -c 700 gpurun_out/bench_$cfg.json; tail -2 gpurun_out/bench_$cfg.err
done
M="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum"
run() { echo "== $1"; shift; timeout 600 ncu $M --clock-control none -k regex:k_dgemm -s 2 -c 1 --csv "$@" 2>&1 | grep -E "k_dgemm" | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' ; }
run "tma default" python tools/gemm_probe.py 4096 16384 16384 0
run "ldgsts" python tools/gemm_probe.py 4096 16384 16384 1
AXB_GEMM_L2PROMO=3 run "tma promo256" python tools/gemm_probe.py 4096 16384 16384 0
AXB_GEMM_L2PROMO=0 run "tma promo none" python tools/gemm_probe.py 4096 16384 16384 0
AXB_GEMM_GROUP=4 run "tma group4" python tools/gemm_probe.py 4096 16384 16384 0
AXB_GEMM_GROUP=16 run "tma group16" python tools/gemm_probe.py 4096 16384 16384 0
AXB_GEMM_GROUP=16 run "ldgsts group16" python tools/gemm_probe.py 4096 16384 16384 1
for g in 4 8 16; do AXB_GEMM_GROUP=$g python tools/gemm_probe.py 4096 16384 16384 0; done
