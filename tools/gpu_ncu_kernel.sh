#!/bin/bash
# usage: gpu_ncu_kernel.sh <tag> <kernel-name-regex> <launch-count> <python script and args...>
# one `ncu --set full` capture of the matching kernels (after the same command has run clean without ncu)
mkdir -p gpurun_out
TAG=$1; RE=$2; CNT=$3; shift 3
python "$@" > gpurun_out/plain_$TAG.log 2>&1 || { cat gpurun_out/plain_$TAG.log; exit 1; }
cat gpurun_out/plain_$TAG.log
ncu --set full --clock-control none --import-source on -k regex:$RE -c $CNT -o gpurun_out/prof_$TAG -f python "$@" > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log; ls -la gpurun_out/prof_$TAG.ncu-rep
