#!/bin/bash
mkdir -p gpurun_out
for tp in 1 3 9; do
echo "== DRB_TRAIN_TAPS=$tp"
DRB_TRAIN_TAPS=$tp python profiles/experiments/train_fwd_probe.py 2>&1 | grep "mask 1:"
rm -f gpurun_out/parity_numbers.log
DRB_TRAIN_TAPS=$tp timeout 600 python -m pytest tests/test_gpu_train.py -q -k "golden or full_frames" 2>&1 | tail -1
grep "train" gpurun_out/parity_numbers.log | cut -c1-175
DRB_TRAIN_TAPS=$tp timeout 600 python profiles/experiments/train_bench.py 16 3 2>&1 | grep -v "sampling loop" | tail -1 | cut -c1-120
done
echo "== DRB_TRAIN_TC=0"
rm -f gpurun_out/parity_numbers.log
DRB_TRAIN_TC=0 timeout 600 python -m pytest tests/test_gpu_train.py -q 2>&1 | tail -1
