#!/bin/bash
# balanced pass-unit ranges in the persistent linear conv (forward conv + dgrad): parity, then A/B timings
mkdir -p gpurun_out; rm -f gpurun_out/parity_numbers.log
timeout 600 python -m pytest tests/test_gpu_train.py -q -x 2>&1 | tail -3
grep "train" gpurun_out/parity_numbers.log | cut -c1-200
for cfg in "0 9" "0 1" "1 9" "1 1"; do
  set -- $cfg
  echo "== DRB_LIN_BALANCE=$1 DRB_TRAIN_DGRAD_TAPS=$2: $(DRB_LIN_BALANCE=$1 DRB_TRAIN_DGRAD_TAPS=$2 timeout 300 python profiles/experiments/train_bench.py 16 5 noeager 2>&1 | grep -v 'sampling loop' | tail -1 | cut -c1-90)"
done
