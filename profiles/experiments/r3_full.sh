#!/bin/bash
# end-of-round checkpoint: full GPU suite, smoke, the driver's two bench commands, ncu launch lists + full captures
mkdir -p gpurun_out
rm -f gpurun_out/parity_numbers.log
timeout 1500 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/pytest_gpu_r3.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_r3.log
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r3.json 2> gpurun_out/bench_r3.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r3_reference.json 2> gpurun_out/bench_r3_reference.err; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/launches_r3.csv \
  python bench.py --lean --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/b_ncu_r3.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r3_train.csv python profiles/experiments/train_prof.py > gpurun_out/r3_train.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r3_clip.csv python profiles/experiments/e2e_prof.py > gpurun_out/r3_clip.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"umma_gate_n4|umma_res_pers|umma_head_pers|umma_zgemm|in_proj" -s 60 -c 8 -f -o gpurun_out/prof_umma_r3 \
  python bench.py --lean --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/prof_r3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"umma_conv_lin_pers|stft_mel|spec_finalize" -c 4 -f -o gpurun_out/prof_clip_r3 \
  python profiles/experiments/e2e_prof.py > gpurun_out/prof_clip_r3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"umma_conv_lin_pers|umma_zgemm|split_pair_T" -s 40 -c 8 -f -o gpurun_out/prof_train_r3 \
  python profiles/experiments/train_prof.py > gpurun_out/prof_train_r3.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
