#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_cuda_parity.py tests/test_widen_cuda.py -q --timeout=200 -k "heaviside or solid or soft_sphere or particle or device_field or misc" 2>&1 | tail -4 | cut -c1-300
timeout 200 python bench.py --config c3 --no-cpu --steps 6 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c3', d['ms_per_step'], d['value'])"
timeout 200 python bench.py --config c3 --reinit --no-cpu --steps 6 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c3+reinit', d['ms_per_step'], d['value'])"
