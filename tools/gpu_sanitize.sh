#!/bin/bash
# compute-sanitizer (memcheck, then racecheck on the shared-memory kernels) over the tests of the round-2 kernels
mkdir -p gpurun_out
T=${TAG:-r02ao}
K="lattice_remesh or fused_solid or heaviside_with_mask or batched_particle_ensemble or soft_sphere_stepper_device or particle_stepper_device or periodic_fft"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 0 python -m pytest tests/test_widen_cuda.py tests/test_cuda_parity.py -q -x -m gpu --timeout=1200 -k "$K" > gpurun_out/${T}_memcheck.txt 2>&1
echo "memcheck rc=$?"; grep -c "Invalid\|out of bounds" gpurun_out/${T}_memcheck.txt; tail -6 gpurun_out/${T}_memcheck.txt | cut -c1-200
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --launch-timeout 0 python -m pytest tests/test_widen_cuda.py tests/test_cuda_parity.py -q -x -m gpu --timeout=800 -k "fused_solid or periodic_fft" > gpurun_out/${T}_racecheck.txt 2>&1
echo "racecheck rc=$?"; grep -c "hazard" gpurun_out/${T}_racecheck.txt; tail -6 gpurun_out/${T}_racecheck.txt | cut -c1-200
