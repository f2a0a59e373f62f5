#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --lean --no-cpu-baseline --steps 20 --warmup 5 2>gpurun_out/bench_r2z.err > gpurun_out/bench_r2z.json
python -c "
import json; l=json.load(open('gpurun_out/bench_r2z.json')); r=l['roofline']
print('bench', round(l['value'],2), 'e2e', round(l['e2e']['value'],2), 'e2e ms/step', round(l['e2e']['ms_per_step'],3), 'traffic', r['traffic'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2z_e2e_launches.csv python profiles/experiments/e2e_prof.py > gpurun_out/r2z_e2e.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2z_train_launches.csv python profiles/experiments/train_prof.py > gpurun_out/r2z_train.log 2>&1
tail -2 gpurun_out/r2z_e2e.log gpurun_out/r2z_train.log
wc -l gpurun_out/r2z_e2e_launches.csv gpurun_out/r2z_train_launches.csv
