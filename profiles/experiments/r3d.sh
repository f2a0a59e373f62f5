#!/bin/bash
# split-K 1x1 weight gradients: training parity + timing (DRB_TRAIN_TC=15: without them)
mkdir -p gpurun_out; rm -f gpurun_out/parity_numbers.log
timeout 600 python -m pytest tests/test_gpu_train.py -q -x 2>&1 | tail -4
grep "train" gpurun_out/parity_numbers.log | cut -c1-200
for e in 15 31; do
  echo "== DRB_TRAIN_TC=$e"
  DRB_TRAIN_TC=$e timeout 300 python profiles/experiments/train_bench.py 16 3 2>&1 | grep -v "sampling loop" | tail -1 | cut -c1-200
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r3d_train_launches.csv python profiles/experiments/train_prof.py > gpurun_out/r3d_train.log 2>&1
tail -n 2 gpurun_out/r3d_train.log
