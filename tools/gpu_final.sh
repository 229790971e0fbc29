#!/bin/bash
# round-end validation on one GPU: parity suite, smoke, the default bench line, the reference arm
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=900 --maxfail=20 > gpurun_out/pytest_final.txt 2>&1
tail -4 gpurun_out/pytest_final.txt | cut -c1-220
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_final_c4.json 2> gpurun_out/bench_final_c4.err; python -c "
import json; d=json.load(open('gpurun_out/bench_final_c4.json')); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['solve_ms'], d['cpu_baseline']['value'], d['e2e']['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks'])"; tail -2 gpurun_out/bench_final_c4.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-300
timeout 200 python bench.py --config c3 --no-cpu --steps 6 --warmup 3 > gpurun_out/bench_final_c3.json 2>/dev/null; timeout 200 python bench.py --config c3 --reinit --no-cpu --steps 6 --warmup 3 > gpurun_out/bench_final_c3_reinit.json 2>/dev/null; timeout 200 python bench.py --config c5 --no-cpu --steps 10 --warmup 3 > gpurun_out/bench_final_c5.json 2>/dev/null
python -c "
import json
for f in ('c3','c3_reinit','c5'):
    d=json.load(open('gpurun_out/bench_final_%s.json'%f)); print(f, d['ms_per_step'], d['value'])"
