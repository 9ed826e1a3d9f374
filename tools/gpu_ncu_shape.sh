#!/bin/bash
# ncu --set full capture of ONE launch of the tcgen05 GEMM kernel on one bench_shapes.py shape:  gpu_ncu_shape.sh <shape-filter> <out-name>
mkdir -p gpurun_out
python tools/bench_shapes.py "$1" --once > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:umma_gemm -c 1 -o gpurun_out/$2 python tools/bench_shapes.py "$1" --once > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log; ls -la gpurun_out/$2.ncu-rep
