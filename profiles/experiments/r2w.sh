#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_numbers.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_chains.py -q -x -k "chain or hoisted or configs1 or forward_vs" 2>&1 | tail -4
grep "configs\[1\]\|chain200\[f16n4\]\|CONDPRE" gpurun_out/parity_numbers.log | cut -c1-200
for e in 0 1; do
DRB_COND_SIMT=$e timeout 300 python bench.py --lean --no-cpu-baseline --steps 20 --warmup 5 2>gpurun_out/bench_r2w_$e.err > gpurun_out/bench_r2w_$e.json
python -c "
import json; l=json.load(open('gpurun_out/bench_r2w_$e.json')); r=l['roofline']
print('COND_SIMT=$e', round(l['value'],2), 'e2e', round(l['e2e']['value'],2), 'e2e ms/step', round(l['e2e']['ms_per_step'],3))"
done
