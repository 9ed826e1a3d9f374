#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run2.log
: > $L
for t in test_in_stats test_layernorm test_patchify test_pwa test_subsample test_attention_linear test_attention_windows test_conv_cin1 test_blend; do
  echo "=== $t" >> $L
  timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -k "$t" 2>&1 | tail -40 >> $L
done
tail -150 $L
