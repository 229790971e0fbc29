#!/bin/bash
# quick A/B of the non-HBM-bound kernels: tests of the touched kernels, c2 / c3 / c5 lines, ncu launch lists
mkdir -p gpurun_out
T=${TAG:-r02aj}
timeout 900 python -m pytest tests/test_widen_cuda.py tests/test_cuda_parity.py -q --timeout=300 -m gpu -k "solid or remesh or p2m or periodic or rfft or soft_sphere or ensemble" 2>&1 | tail -6 | cut -c1-300
for c in c2 c3 c5; do
  timeout 300 python bench.py --config $c --no-cpu --steps 20 --warmup 3 > gpurun_out/${T}_bench_$c.json 2> gpurun_out/${T}_bench_$c.err
done
for c in c3d c5b c2; do
  timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_$c.csv python tools/profile_config.py $c 3 > gpurun_out/${T}_launches_$c.log 2>&1
  python tools/launch_summary.py gpurun_out/${T}_launches_$c.csv 50 2>&1 | grep -v "at::" > gpurun_out/${T}_kernel_summary_$c.txt
done
python - <<PY
import json
for c in ("c2", "c3", "c5"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/${T}_bench_{c}.json") if l.startswith("{")][-1])
        print(c, d["ms_per_step"], d["value"], d["roofline"]["frac"], d["e2e"]["ms_per_step"], d.get("step_roofline", {}).get("frac"), d["gpu_launches"])
    except Exception as e:
        print(c, "failed", e); print(open(f"gpurun_out/${T}_bench_{c}.err").read()[-1500:])
PY
grep -h "k_solid_fused\|k_p2m_lattice_gather\|k_rfft_rows\|k_irfft_rows\|k_bubble_pre" gpurun_out/${T}_kernel_summary_*.txt
