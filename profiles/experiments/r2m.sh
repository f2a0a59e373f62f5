#!/bin/bash
mkdir -p gpurun_out
python profiles/experiments/batch_indep2.py 2>&1 | grep -v "sampling loop" | tail -30
timeout 300 ncu --set full --clock-control none --import-source on -k regex:umma_gate_n4 -s 20 -c 2 -f -o gpurun_out/prof_n4d \
  python bench.py --lean --no-cpu-baseline --precision f16n4 --steps 5 --warmup 3 > gpurun_out/prof_n4d.log 2>&1
