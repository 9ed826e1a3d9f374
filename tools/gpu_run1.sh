#!/bin/bash
# first GPU contact: each test function in its own process so a trap in one does not poison the rest
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/run1.log 2>&1
for t in test_linear_bf16 test_linear_bias test_head test_conv3x3x3 test_conv_transpose test_concat; do
  echo "=== $t" >> gpurun_out/run1.log
  timeout 600 python -m pytest tests/test_umma_gemm_gpu.py -m gpu -q --tb=short -k "$t" >> gpurun_out/run1.log 2>&1
  echo "exit $?" >> gpurun_out/run1.log
done
tail -100 gpurun_out/run1.log
