#!/bin/bash
# round-2 checkpoint run: full GPU suite, smoke, the driver's bench command, ncu launch list and full captures
mkdir -p gpurun_out
rm -f gpurun_out/parity_numbers.log
timeout 1200 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu_r2d.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_r2d.log
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; echo "bench rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/launches_r2d.csv \
  python bench.py --lean --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/b_ncu_r2d.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"umma_gate_n4|umma_res_pers|umma_zgemm|simt_gemm|in_proj" -s 60 -c 8 -f -o gpurun_out/prof_r2d \
  python bench.py --lean --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/prof_r2d.log 2>&1
ls -la gpurun_out/prof_r2d.ncu-rep
