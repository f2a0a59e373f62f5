#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_numbers.log
timeout 900 python -m pytest tests/test_gpu_train.py -q -x 2>&1 | tail -25
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_chains.py -q -x -k "chain or hoisted or per_layer" 2>&1 | tail -5
timeout 300 python bench.py --lean --no-cpu-baseline --steps 20 --warmup 5 2>gpurun_out/bench_r2e.err > gpurun_out/bench_r2e.json
python -c "
import json; l=json.load(open('gpurun_out/bench_r2e.json')); r=l['roofline']
print('value', round(l['value'],2), 'e2e', round(l['e2e']['value'],2), {k: round(v,3) for k,v in r['per_step_ms'].items()})"
