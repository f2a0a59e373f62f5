#!/bin/bash
mkdir -p gpurun_out
for m in 0 1 2 8; do
rm -f gpurun_out/parity_numbers.log
DRB_TRAIN_TC=$m timeout 300 python -m pytest tests/test_gpu_train.py -q -k ragged 2>&1 | tail -1
echo "mask $m: $(grep ragged gpurun_out/parity_numbers.log)"
done
