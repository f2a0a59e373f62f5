#!/bin/bash
mkdir -p gpurun_out
timeout 300 python profiles/experiments/ragged_probe.py 2>&1 | grep -v "sampling loop" | tail -12
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r3b_e2e_launches.csv python profiles/experiments/e2e_prof.py > gpurun_out/r3b_e2e.log 2>&1
tail -n 2 gpurun_out/r3b_e2e.log
