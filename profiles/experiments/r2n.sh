#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_numbers.log
timeout 900 python -m pytest tests -m gpu -x -q -k "f16n4 or loop_equals" 2>&1 | tail -4
timeout 300 python bench.py --lean --no-cpu-baseline --steps 50 --warmup 5 2>gpurun_out/bench_r2n.err > gpurun_out/bench_r2n.json
python -c "
import json; l=json.load(open('gpurun_out/bench_r2n.json')); r=l['roofline']
print('balanced-last-round', round(l['value'],2), 'e2e', round(l['e2e']['value'],2), 'gate_full_ms', round(r['full_launch_avg_ms'],4), 'frac', round(r['frac'],4), {k: round(v,3) for k,v in r['per_step_ms'].items()}, l['clocks'])"
timeout 300 python bench.py --lean --no-cpu-baseline --steps 50 --warmup 5 --config 3 --batch 8 2>/dev/null | python -c "
import json,sys; l=json.loads(sys.stdin.readline()); r=l['roofline']
print('8-roll shard (configs[3] at N=4):', round(l['value'],2), 'steps/s', {k: round(v,3) for k,v in r['per_step_ms'].items()})"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:umma_gate_n4 -s 20 -c 2 -f -o gpurun_out/prof_n4e \
  python bench.py --lean --no-cpu-baseline --precision f16n4 --steps 5 --warmup 3 > gpurun_out/prof_n4e.log 2>&1
