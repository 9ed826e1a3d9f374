#!/bin/bash
# usage: gpu_ncu_full.sh <tag> <shape-filter...>   one `ncu --set full` capture of the filtered shapes (1 launch each)
mkdir -p gpurun_out
TAG=$1; shift
python tools/bench_shapes.py "$@" > gpurun_out/shapes_$TAG.log 2>&1; cat gpurun_out/shapes_$TAG.log
python tools/bench_shapes.py "$@" --once > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:umma_gemm -o gpurun_out/prof_$TAG -f python tools/bench_shapes.py "$@" --once > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log; ls -la gpurun_out/prof_$TAG.ncu-rep
