#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_numbers.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_valstep.py tests/test_gpu_eager_baseline.py tests/test_gpu_full_chains.py -q -x 2>&1 | tail -12
for e in 0 1; do
DRB_NO_HEAD_TC=$e timeout 300 python bench.py --lean --no-cpu-baseline --steps 50 --warmup 5 2>gpurun_out/bench_r2h_$e.err > gpurun_out/bench_r2h_$e.json
python -c "
import json; l=json.load(open('gpurun_out/bench_r2h_$e.json')); r=l['roofline']
print('NO_HEAD_TC=$e', round(l['value'],2), 'e2e', round(l['e2e']['value'],2), {k: round(v,3) for k,v in r['per_step_ms'].items()}, l['clocks'])"
done
