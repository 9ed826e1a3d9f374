#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run3.log
: > $L
echo "=== blocks" >> $L
timeout 900 python -m pytest tests/test_blocks_gpu.py -m gpu -q --tb=short 2>&1 | tail -60 >> $L
echo "=== models" >> $L
timeout 1200 python -m pytest tests/test_models_gpu.py -m gpu -q --tb=short -s 2>&1 | tail -120 >> $L
tail -200 $L
