#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:umma_zgemm -s 8 -c 4 -f -o gpurun_out/prof_head \
  python bench.py --lean --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/prof_head.log 2>&1
