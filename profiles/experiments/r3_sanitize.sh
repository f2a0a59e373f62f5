#!/bin/bash
# compute-sanitizer memcheck + synccheck over the kernels added in the second half of round 2: umma_conv_lin_pers_kernel (+ lin_fixup_kernel at
# 10 rolls), split-K zgemm + splitk_reduce, skinny_gemm, adam_multi, stft_mel_kernel.  Bounded runs.
mkdir -p gpurun_out
echo "=== memcheck: training step (goldens B=2 x 128) + fused mel" > gpurun_out/r3_sanitize_memcheck.log
timeout 280 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_train.py tests/test_gpu_parity.py -q -x -k "(golden and x_0_l2) or mel_frontend or adam_step" >> gpurun_out/r3_sanitize_memcheck.log 2>&1
echo "rc=$?" >> gpurun_out/r3_sanitize_memcheck.log
tail -4 gpurun_out/r3_sanitize_memcheck.log
echo "=== memcheck: 10 rolls x 640 frames (pass-range split, parked partial tile + fix-up)" > gpurun_out/r3_sanitize_memcheck_b10.log
timeout 280 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_train.py -q -x -k "do_not_divide and 10" >> gpurun_out/r3_sanitize_memcheck_b10.log 2>&1
echo "rc=$?" >> gpurun_out/r3_sanitize_memcheck_b10.log
tail -4 gpurun_out/r3_sanitize_memcheck_b10.log
echo "=== synccheck: training step (goldens B=2 x 128) + fused mel" > gpurun_out/r3_sanitize_synccheck.log
timeout 200 compute-sanitizer --tool synccheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_train.py tests/test_gpu_parity.py -q -x -k "(golden and x_0_l2) or mel_frontend" >> gpurun_out/r3_sanitize_synccheck.log 2>&1
echo "rc=$?" >> gpurun_out/r3_sanitize_synccheck.log
tail -4 gpurun_out/r3_sanitize_synccheck.log
