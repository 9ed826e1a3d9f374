#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_umma_gemm_gpu.py -m gpu -q --tb=short 2>&1 | tail -5
for v in 0 3; do
  echo "=== variant $v"
  CTU_GEMM_VARIANT=$v python tools/bench_shapes.py 2>&1 | tail -20
done
