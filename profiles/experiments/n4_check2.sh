#!/bin/bash
# f16n4 parity tests, bench, and an ncu capture of the n4 kernel in one call
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "f16n4" 2>&1 | tail -4
timeout 300 python bench.py --lean --no-cpu-baseline --precision f16n4 --steps 50 --warmup 5 2>gpurun_out/bench_r2g.err > gpurun_out/bench_r2g.json
python -c "
import json; l=json.load(open('gpurun_out/bench_r2g.json')); r=l['roofline']
print('f16n4', round(l['value'],2), 'e2e', round(l['e2e']['value'],2), 'gate_full_ms', round(r['full_launch_avg_ms'],4), 'frac', round(r['frac'],4), {k: round(v,3) for k,v in r['per_step_ms'].items()}, l['clocks'])"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:umma_gate_n4 -s 20 -c 2 -f -o gpurun_out/prof_n4c \
  python bench.py --lean --no-cpu-baseline --precision f16n4 --steps 5 --warmup 3 > gpurun_out/prof_n4c.log 2>&1
