#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_numbers.log
timeout 900 python -m pytest tests/test_gpu_dist.py -q -rs 2>&1 | grep -v "sampling loop" | tail -40 | tee gpurun_out/pytest_gpu_dist_2gpu.log
cp gpurun_out/parity_numbers.log gpurun_out/parity_numbers_dist.log 2>/dev/null
