#!/bin/bash
mkdir -p gpurun_out
python profiles/experiments/train_fwd_probe.py 2>&1 | grep -v "sampling loop" | tail -4
rm -f gpurun_out/parity_numbers.log
timeout 600 python -m pytest tests/test_gpu_train.py -q 2>&1 | tail -3
grep "train" gpurun_out/parity_numbers.log | cut -c1-220
timeout 600 python profiles/experiments/train_bench.py 16 3 2>&1 | grep -v "sampling loop" | tail -1 | tee gpurun_out/train_bench_tc.json
