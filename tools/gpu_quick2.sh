#!/bin/bash
# 1 GPU: the new tests + the physics pin + the whole gpu suite
mkdir -p gpurun_out
TAG=${1:-r02n}
timeout 900 python -m pytest tests/test_physics_cuda.py -x -q -m gpu -s > gpurun_out/${TAG}_physics.txt 2>&1
tail -4 gpurun_out/${TAG}_physics.txt
timeout 1700 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest_gpu.txt 2>&1
tail -4 gpurun_out/${TAG}_pytest_gpu.txt
