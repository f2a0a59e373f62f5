#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_numbers.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_chains.py -q -x -k "per_layer or forward_vs or chain_transcription or sampler_single or configs1 or ragged or per_sample" 2>&1 | tail -6
for e in 0 1; do
DRB_NO_HEAD_TC=$e timeout 300 python bench.py --lean --no-cpu-baseline --steps 50 --warmup 5 2>gpurun_out/bench_r2i_$e.err > gpurun_out/bench_r2i_$e.json
python -c "
import json; l=json.load(open('gpurun_out/bench_r2i_$e.json')); r=l['roofline']
print('NO_HEAD_TC=$e', round(l['value'],2), 'e2e', round(l['e2e']['value'],2), {k: round(v,3) for k,v in r['per_step_ms'].items()}, l['clocks'])"
done
