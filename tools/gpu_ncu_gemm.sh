#!/bin/bash
mkdir -p gpurun_out
python tools/bench_shapes.py ffn128_up_noact --once > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:umma_gemm -c 2 -o gpurun_out/prof_ffn128up python tools/bench_shapes.py ffn128_up_noact --once > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log; ls -la gpurun_out/prof_ffn128up.ncu-rep
