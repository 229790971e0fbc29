#!/bin/bash
# round-2 evidence on one GPU: full GPU test suite, smoke, the default bench line (with the c1/c2/c3/c5 objects), the
# reference arm, the ncu launch list of the C4 step and the `--set full` capture of its solve kernels
mkdir -p gpurun_out
T=${TAG:-r02}
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 --maxfail=20 > gpurun_out/${T}_pytest_gpu.txt 2>&1
tail -4 gpurun_out/${T}_pytest_gpu.txt | cut -c1-220
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python bench.py > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err
tail -2 gpurun_out/${T}_bench_1gpu.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_bench_reference_arm.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_c4.csv python tools/profile_step.py 16384 3 > gpurun_out/${T}_launches_c4.log 2>&1
python tools/launch_summary.py gpurun_out/${T}_launches_c4.csv 40 2>&1 | grep -v "at::" | head -40 > gpurun_out/${T}_kernel_summary_c4.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_dct_rows_w|k_tri_sweep_ws|k_tri_sweep_tma" -c 8 -o gpurun_out/${T}_ncu_full_solve_c4 -f python tools/profile_step.py 16384 2 > gpurun_out/${T}_ncu_full_solve_c4.log 2>&1
python tools/ncu_summary.py gpurun_out/${T}_ncu_full_solve_c4.ncu-rep > gpurun_out/${T}_ncu_full_solve_c4_summary.csv 2>/dev/null
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/${T}_bench_1gpu.json") if l.startswith("{")][-1])
print("c4", d["ms_per_step"], d["value"], d["roofline"]["frac"], d["roofline"]["solve_ms"], (d.get("cpu_baseline") or {}).get("value"), d["e2e"]["ms_per_step"], d["gpu_launches"], d["clocks"])
for k, c in (d.get("configs") or {}).items():
    print(k, c.get("ms_per_step"), c.get("value"), (c.get("roofline") or {}).get("frac"), (c.get("cpu_baseline") or {}).get("value"), (c.get("e2e") or {}).get("ms_per_step"))
PY
cut -c1-400 gpurun_out/${T}_bench_reference_arm.json
cat gpurun_out/${T}_kernel_summary_c4.txt
# launch lists of the other configurations' steps (eager launches of the device-resident steppers)
for c in c2 c3d c5b; do
  timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_$c.csv python tools/profile_config.py $c 3 > gpurun_out/${T}_launches_$c.log 2>&1
  python tools/launch_summary.py gpurun_out/${T}_launches_$c.csv 40 2>&1 | grep -v "at::" | head -36 > gpurun_out/${T}_kernel_summary_$c.txt
done
head -24 gpurun_out/${T}_kernel_summary_c3d.txt
