#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run7.log
: > $L
for t in test_linear_bf16 test_linear_bias test_head test_conv3x3x3 test_conv_transpose test_concat; do
  echo "=== $t" >> $L
  timeout 600 python -m pytest tests/test_umma_gemm_gpu.py -m gpu -q --tb=short -k "$t" 2>&1 | tail -25 >> $L
done
echo "=== blocks+models+sw" >> $L
timeout 1500 python -m pytest tests/test_blocks_gpu.py tests/test_models_gpu.py tests/test_sliding_window_gpu.py -m gpu -q --tb=short 2>&1 | tail -30 >> $L
echo "=== timing" >> $L
timeout 600 python tools/time_forward.py >> $L 2>&1
tail -100 $L
