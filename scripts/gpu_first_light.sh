#!/bin/bash
# First-light on the B200 box: each kernel family in its own process (a device trap is sticky).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for k in test_base_gemm_and_T_fp32_out test_fused_forward_matches_oracle test_backward_kernels_match_oracle test_linearity_full_size test_rejects_bad_arguments; do
  echo "=== $k" >> gpurun_out/first_light.log
  timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "$k" --timeout 120 -p no:cacheprovider 2>&1 | tail -40 >> gpurun_out/first_light.log
done
tail -5 gpurun_out/first_light.log
