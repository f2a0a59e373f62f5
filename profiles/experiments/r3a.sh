#!/bin/bash
# persistent linear conv kernel (umma_conv_lin_pers_kernel): training parity, cond-table parity, A/B timings
mkdir -p gpurun_out; rm -f gpurun_out/parity_numbers.log
timeout 600 python -m pytest tests/test_gpu_train.py -q -x 2>&1 | tail -4
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_chains.py -q -x -k "chain_transcription_200 or configs1_full_chain" 2>&1 | tail -3
grep "train\|chain200\[f16n4\]\|configs\[1\]" gpurun_out/parity_numbers.log | cut -c1-180
for e in 0 1; do
  echo "== DRB_LIN_PERS=$e"
  DRB_LIN_PERS=$e timeout 300 python profiles/experiments/train_bench.py 16 3 2>&1 | grep -v "sampling loop" | tail -1 | cut -c1-200
  DRB_LIN_PERS=$e timeout 300 python bench.py --lean --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys; l=json.loads(sys.stdin.read()); print('bench', round(l['value'],2), 'e2e', round(l['e2e']['value'],2), 'e2e ms/step', round(l['e2e']['ms_per_step'],3))"
done
