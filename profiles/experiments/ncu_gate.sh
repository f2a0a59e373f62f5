#!/bin/bash
# one ncu --set full capture (2 launches each) of the f16n4 and f16e5 gate kernels inside a short bench run
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:umma_gate_n4 -s 20 -c 2 -f -o gpurun_out/prof_n4 \
  python bench.py --lean --no-cpu-baseline --precision f16n4 --steps 5 --warmup 3 > gpurun_out/prof_n4.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:umma_gate_pers -s 20 -c 2 -f -o gpurun_out/prof_e5 \
  python bench.py --lean --no-cpu-baseline --precision f16e5 --steps 5 --warmup 3 > gpurun_out/prof_e5.log 2>&1
tail -3 gpurun_out/prof_n4.log gpurun_out/prof_e5.log
