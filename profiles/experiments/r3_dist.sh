#!/bin/bash
# final state on 2 GPUs: N-rank == 1-rank tests over NCCL, then the driver's 2-GPU bench command (lean)
mkdir -p gpurun_out; rm -f gpurun_out/parity_numbers.log
timeout 900 python -m pytest tests/test_gpu_dist.py -q -rs 2>&1 | grep -v "sampling loop" | tail -12 | tee gpurun_out/r3_pytest_gpu_dist_2gpu.log
cp gpurun_out/parity_numbers.log gpurun_out/r3_parity_numbers_dist.log 2>/dev/null
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --lean --no-cpu-baseline > gpurun_out/bench_r3_2gpu.json 2> gpurun_out/bench_r3_2gpu.err; echo "bench rc=$?"
python -c "
import json; l=json.load(open('gpurun_out/bench_r3_2gpu.json')); print('2 GPUs:', round(l['value'],2), 'e2e', round(l['e2e']['value'],2), l['e2e']['ms_per_call_all'], l['scaling'])"
