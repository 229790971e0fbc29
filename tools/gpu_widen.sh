#!/bin/bash
# SURVEY 8f rows on one GPU: the new parity tests first, then smoke, the C3 step with / without the level-set
# re-initialisation, the default bench line, and as much of the full GPU suite as the time allows.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_widen_cuda.py -q -s --timeout=300 > gpurun_out/pytest_widen.txt 2>&1
tail -25 gpurun_out/pytest_widen.txt | cut -c1-200
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python tools/bench_widen.py 2>&1 | tail -6 | cut -c1-250
timeout 200 python bench.py --config c3 --reinit --no-cpu --steps 6 --warmup 3 > gpurun_out/bench_c3_reinit.json 2> gpurun_out/bench_c3_reinit.err
timeout 200 python bench.py --config c3 --no-cpu --steps 6 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
python - <<'PY'
import json
for f in ("bench_c3_reinit", "bench_c3"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json")); print(f, d["ms_per_step"], d["value"], d["config"]["workload"][:90])
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/{f}.err").read()[-600:])
PY
timeout 200 python -m pytest tests/test_cuda_parity.py -q --timeout=180 -k "soft_sphere or device_field or particle" > gpurun_out/pytest_steppers.txt 2>&1
tail -3 gpurun_out/pytest_steppers.txt | cut -c1-200
timeout 240 python bench.py --no-cpu > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
python -c "
import json; d=json.load(open('gpurun_out/bench_c4.json')); print('c4', d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['gpu_launches'])" || tail -5 gpurun_out/bench_c4.err
timeout 420 python -m pytest tests -m gpu -q --timeout=300 --maxfail=10 -x --deselect tests/test_widen_cuda.py > gpurun_out/pytest_full.txt 2>&1
tail -3 gpurun_out/pytest_full.txt | cut -c1-200
