#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=900 --maxfail=30 -k "stepper or glue" > gpurun_out/pytest_gpu7.txt 2>&1
tail -60 gpurun_out/pytest_gpu7.txt | cut -c1-200
