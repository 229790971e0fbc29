#!/bin/bash
# edit-measure loop of the 8f kernels: their parity tests, the micro-benchmark, the C3 step with the re-initialisation,
# and the ncu launch list of one re-initialisation
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_widen_cuda.py -q -s --timeout=300 > gpurun_out/pytest_widen.txt 2>&1
tail -25 gpurun_out/pytest_widen.txt | cut -c1-200
timeout 200 python tools/bench_widen.py 2>&1 | tail -6 | cut -c1-250
timeout 200 python bench.py --config c3 --reinit --no-cpu --steps 6 --warmup 3 > gpurun_out/bench_c3_reinit.json 2> gpurun_out/bench_c3_reinit.err
python -c "
import json; d=json.load(open('gpurun_out/bench_c3_reinit.json')); print('c3+reinit', d['ms_per_step'], d['value'])" || tail -5 gpurun_out/bench_c3_reinit.err
./tools/gpu_profile_reinit.sh 2>&1 | grep -v "at::"
grep "k_reinit_sweep" gpurun_out/launches_reinit.csv | grep "gpu__time_duration" | awk -F'","' '{printf "%s ", $NF}'; echo
