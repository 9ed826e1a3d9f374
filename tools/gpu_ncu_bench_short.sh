#!/bin/bash
# ncu launch list of bench.py itself, first ~2 training steps only (the steps of the warm-up: same kernels as the replayed graph)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-sliding-window --no-cpu-baseline"
$CMD > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3800 --csv --log-file gpurun_out/launches_bench_v3.csv $CMD > gpurun_out/bench_under_ncu.json 2> gpurun_out/ncu_bench.err
cut -c1-200 gpurun_out/bench_plain.json
wc -l gpurun_out/launches_bench_v3.csv
