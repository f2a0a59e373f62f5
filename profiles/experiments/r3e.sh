#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_numbers.log
timeout 600 python -m pytest tests/test_gpu_train.py -q -x 2>&1 | tail -3
grep "train" gpurun_out/parity_numbers.log | cut -c1-200
timeout 300 python profiles/experiments/train_bench.py 16 5 2>&1 | grep -v "sampling loop" | tail -1 | cut -c1-200
