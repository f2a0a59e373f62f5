#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_numbers.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "mel or ragged or silent" 2>&1 | tail -2
grep "mel" gpurun_out/parity_numbers.log | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r3h_e2e_launches.csv python profiles/experiments/e2e_prof.py > gpurun_out/r3h_e2e.log 2>&1
grep "stft_mel\|spec_finalize\|minmax\|conv_lin" gpurun_out/r3h_e2e_launches.csv | awk -F'","' '{print $5, $NF}' | sed 's/(.*)//; s/"//g' | sort | uniq -c
