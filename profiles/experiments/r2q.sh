#!/bin/bash
mkdir -p gpurun_out
for m in 1 2 3; do
rm -f gpurun_out/parity_numbers.log
DRB_TRAIN_TC=$m timeout 600 python -m pytest tests/test_gpu_train.py -q 2>&1 | tail -3
echo "== DRB_TRAIN_TC=$m"; grep "train" gpurun_out/parity_numbers.log | cut -c1-230
done
