#!/bin/bash
# compute-sanitizer over the tests that exercise all five UMMA kernels (gate n4 / pers / win, RES pers, zgemm HEAD) and the
# SIMT epilogues: memcheck, synccheck, racecheck.  Each run is bounded; a tool that cannot instrument tcgen05 / TMA code says so in its log.
mkdir -p gpurun_out
SEL='test_forward_vs_golden or test_tensor_path_matches_fp32_path_per_layer or test_ragged_frames_and_trim or test_forward_per_sample_diffusion_steps or test_loop_equals_repeated_steps_bitwise'
for tool in memcheck synccheck racecheck; do
  echo "=== $tool" > gpurun_out/sanitize_$tool.log
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -x -k "$SEL" >> gpurun_out/sanitize_$tool.log 2>&1
  echo "rc=$?" >> gpurun_out/sanitize_$tool.log
  tail -6 gpurun_out/sanitize_$tool.log
done
