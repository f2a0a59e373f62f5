#!/bin/bash
# the driver's exact 2-GPU command (all extras), final state
mkdir -p gpurun_out
SECONDS=0
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r3_2gpu_full.json 2> gpurun_out/bench_r3_2gpu_full.err; echo "bench rc=$? wall ${SECONDS}s"
python -c "
import json; l=json.load(open('gpurun_out/bench_r3_2gpu_full.json')); print('2 GPUs:', round(l['value'],2), 'e2e', round(l['e2e']['value'],2), l['e2e']['ms_per_call_all'], 'keys', sorted(k for k in l if k in ('configs2','strong','gpu_eager_baseline','train_step','cpu_baseline')))
print('strong', {k:(round(v['value'],2), round(v['e2e']['value'],2)) for k,v in l['strong'].items()})"
tail -3 gpurun_out/bench_r3_2gpu_full.err
