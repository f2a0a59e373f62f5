#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_numbers.log
timeout 600 python -m pytest tests/test_gpu_train.py -q 2>&1 | tail -4
grep "ragged" gpurun_out/parity_numbers.log
echo "=== memcheck (training step kernels)" > gpurun_out/sanitize_train_memcheck.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_train.py -q -x -k "golden and x_0 or ragged" >> gpurun_out/sanitize_train_memcheck.log 2>&1
echo "rc=$?" >> gpurun_out/sanitize_train_memcheck.log
tail -5 gpurun_out/sanitize_train_memcheck.log
