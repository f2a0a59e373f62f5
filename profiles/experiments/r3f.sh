#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_numbers.log
timeout 600 python -m pytest tests/test_gpu_train.py -q -x 2>&1 | tail -3
grep "adam\|3 updates\|Adam" gpurun_out/parity_numbers.log | cut -c1-200
timeout 300 python profiles/experiments/train_bench.py 16 5 2>&1 | grep -v "sampling loop" | tail -1 | cut -c1-200
timeout 300 python bench.py --lean --no-cpu-baseline --steps 20 --warmup 5 2>gpurun_out/bench_r3f.err > gpurun_out/bench_r3f.json
python -c "
import json; l=json.load(open('gpurun_out/bench_r3f.json')); r=l['roofline']
print('bench', round(l['value'],2), 'e2e', round(l['e2e']['value'],2), l['e2e']['ms_per_call_all'], 'frac', round(r['frac'],3))"
