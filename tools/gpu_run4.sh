#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run4.log
: > $L
echo "=== blocks+models" >> $L
timeout 1200 python -m pytest tests/test_blocks_gpu.py tests/test_models_gpu.py -m gpu -q --tb=short 2>&1 | tail -30 >> $L
echo "=== sliding window" >> $L
timeout 1200 python -m pytest tests/test_sliding_window_gpu.py -m gpu -q --tb=short 2>&1 | tail -40 >> $L
echo "=== timing" >> $L
timeout 600 python tools/time_forward.py >> $L 2>&1
tail -120 $L
