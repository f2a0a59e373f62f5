#!/bin/bash
mkdir -p gpurun_out
for m in 7 14; do
rm -f gpurun_out/parity_numbers.log
DRB_TRAIN_TC=$m timeout 600 python -m pytest tests/test_gpu_train.py -q -k "golden or full_frames" 2>&1 | tail -2
echo "== DRB_TRAIN_TC=$m"; grep "train" gpurun_out/parity_numbers.log | cut -c1-200
done
