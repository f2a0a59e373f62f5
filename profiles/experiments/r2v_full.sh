#!/bin/bash
# end-of-round checkpoint: full GPU suite, smoke, the driver's two bench commands, ncu launch list + full captures of one step
mkdir -p gpurun_out
rm -f gpurun_out/parity_numbers.log
timeout 1500 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/pytest_gpu_r2v.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_r2v.log
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r2v.json 2> gpurun_out/bench_r2v.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r2v_reference.json 2> gpurun_out/bench_r2v_reference.err; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/launches_r2v.csv \
  python bench.py --lean --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/b_ncu_r2v.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"umma_gate_n4|umma_res_pers|umma_head_pers|umma_zgemm|in_proj" -s 60 -c 8 -f -o gpurun_out/prof_umma_r2v \
  python bench.py --lean --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/prof_r2v.log 2>&1
ls -la gpurun_out/prof_umma_r2v.ncu-rep
