#!/bin/bash
mkdir -p gpurun_out
python tools/one_conv.py > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv3_halo -s 1 -c 2 -o gpurun_out/prof_halo64 python tools/one_conv.py > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/prof_halo64.ncu-rep
