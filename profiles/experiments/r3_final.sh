#!/bin/bash
# last checkpoint of the round: full GPU suite on the final tree, then bench.py with NO flags (defaults: 1 GPU, 200 steps) for its wall time
mkdir -p gpurun_out; rm -f gpurun_out/parity_numbers.log
timeout 1200 python -m pytest tests -m gpu -q --durations=4 > gpurun_out/pytest_gpu_r3_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r3_final.log
SECONDS=0
timeout 900 python bench.py > gpurun_out/bench_r3_defaults.json 2> gpurun_out/bench_r3_defaults.err; echo "bench (no flags) rc=$? wall ${SECONDS}s"
python -c "
import json; l=json.load(open('gpurun_out/bench_r3_defaults.json')); print('defaults:', 'steps', l['steps'], 'value', round(l['value'],2), 'e2e', round(l['e2e']['value'],2), l['e2e']['ms_per_call_all'], 'frac', round(l['roofline']['frac'],3), 'train', round(l['train_step']['ours_ms'],2))"
