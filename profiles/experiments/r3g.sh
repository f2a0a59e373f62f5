#!/bin/bash
# fused STFT + mel kernel: parity and timing (DRB_MEL_CUFFT=1: the cuFFT pipeline)
mkdir -p gpurun_out; rm -f gpurun_out/parity_numbers.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "mel or forward_vs or ragged or silent or chain_transcription_200" 2>&1 | tail -3
grep "mel\|chain200\[f16n4\]" gpurun_out/parity_numbers.log | cut -c1-200
for e in 1 0; do
DRB_MEL_CUFFT=$e timeout 300 python bench.py --lean --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys; l=json.loads(sys.stdin.read()); m=l['roofline_hbm']['mel']
print('DRB_MEL_CUFFT=$e bench', round(l['value'],2), 'e2e', round(l['e2e']['value'],2), l['e2e']['ms_per_call_all'], 'mel us', round(m['us'],1), 'frac', round(m['frac'],3), 'ws', l['workspace_bytes'])"
done
