#!/bin/bash
# n4 emission with the part multipliers folded into the block scale: the parity numbers must not move (bit-identical codes); timing
mkdir -p gpurun_out; rm -f gpurun_out/parity_numbers.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_chains.py -q -x -k "chain_transcription_200 or forward_vs or configs1_full_chain" 2>&1 | tail -2
grep "configs\[1\] transcription chain \[f16n4\]\|chain200\[f16n4\]\|forward\[f16n4\] max" gpurun_out/parity_numbers.log | cut -c1-200
timeout 300 python bench.py --lean --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys; l=json.loads(sys.stdin.read()); r=l['roofline']
print('bench', round(l['value'],2), 'e2e', round(l['e2e']['value'],2), {k:round(v,3) for k,v in r['per_step_ms'].items()})"
