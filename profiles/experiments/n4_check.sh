#!/bin/bash
# f16n4 parity tests, then the f16n4 / f16e5 bench back to back on the same box
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "f16n4" 2>&1 | tail -15
for p in f16n4 f16e5; do
  timeout 300 python bench.py --lean --no-cpu-baseline --precision $p --steps 50 --warmup 5 2>gpurun_out/bench_r2c_$p.err > gpurun_out/bench_r2c_$p.json
  python -c "
import json; l=json.load(open('gpurun_out/bench_r2c_$p.json')); r=l['roofline']
print('$p', round(l['value'],2), 'e2e', round(l['e2e']['value'],2), 'gate_full_ms', round(r['full_launch_avg_ms'],4), 'frac', round(r['frac'],4), {k: round(v,3) for k,v in r['per_step_ms'].items()}, l['clocks'])"
done
