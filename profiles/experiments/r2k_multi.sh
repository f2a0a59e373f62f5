#!/bin/bash
# multi-GPU evidence: N-rank == 1-rank parity tests over NCCL, then BASELINE configs[3] (32 rolls over 4 GPUs) and configs[1] at N=4
mkdir -p gpurun_out; rm -f gpurun_out/parity_numbers.log
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_dist.py -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_dist_4gpu.log
cp gpurun_out/parity_numbers.log gpurun_out/parity_numbers_dist.log 2>/dev/null
for cfg in 3 1; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 50 --warmup 5 --config $cfg --lean --no-cpu-baseline > gpurun_out/bench_r2k_cfg${cfg}_4gpu.json 2> gpurun_out/bench_r2k_cfg${cfg}_4gpu.err
python -c "
import json; l=json.load(open('gpurun_out/bench_r2k_cfg${cfg}_4gpu.json'))
print('config $cfg N=4', round(l['value'],2), 'e2e', round(l['e2e']['value'],2), l['scaling'], l['config']['workload'][:80])"
done
