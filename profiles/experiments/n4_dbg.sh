#!/bin/bash
# f16n4 gate kernel bottleneck isolation: DRB_N4_DBG bits switch parts of the MMA thread's work off (results are wrong, timing only)
mkdir -p gpurun_out
for d in 0 1 2 3 4 8 12 15; do
  DRB_N4_DBG=$d python bench.py --lean --no-cpu-baseline --precision f16n4 --steps 20 --warmup 3 2>gpurun_out/n4dbg_$d.err | \
    python -c "import sys,json; l=json.loads(sys.stdin.readline()); r=l['roofline']; print('dbg=$d', 'value', round(l['value'],2), 'gate_full_ms', round(r['full_launch_avg_ms'],4), 'per_step', {k: round(v,3) for k,v in r['per_step_ms'].items()}, l['clocks'])"
done | tee gpurun_out/n4dbg.log
python bench.py --lean --no-cpu-baseline --precision f16e5 --steps 20 --warmup 3 2>/dev/null | python -c "import sys,json; l=json.loads(sys.stdin.readline()); r=l['roofline']; print('f16e5', 'value', round(l['value'],2), 'gate_full_ms', round(r['full_launch_avg_ms'],4), 'per_step', {k: round(v,3) for k,v in r['per_step_ms'].items()}, l['clocks'])" | tee -a gpurun_out/n4dbg.log
