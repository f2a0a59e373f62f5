#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:umma_gate_n4 -s 20 -c 2 -f -o gpurun_out/prof_n4b \
  python bench.py --lean --no-cpu-baseline --precision f16n4 --steps 5 --warmup 3 > gpurun_out/prof_n4b.log 2>&1
nvidia-smi --query-gpu=power.limit,power.max_limit,power.default_limit,clocks.max.sm --format=csv
