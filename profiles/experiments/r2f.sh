#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -q -x 2>&1 | tail -5
timeout 600 python profiles/experiments/train_bench.py 16 3 2>/dev/null | tee gpurun_out/train_bench.json
