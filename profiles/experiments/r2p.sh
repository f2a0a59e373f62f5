#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_numbers.log
timeout 600 python -m pytest tests/test_gpu_train.py -q -x 2>&1 | tail -6
grep "train" gpurun_out/parity_numbers.log
timeout 600 python profiles/experiments/train_bench.py 16 3 2>&1 | grep -v "sampling loop" | tail -3 | tee gpurun_out/train_bench_tc.json
