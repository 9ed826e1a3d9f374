#!/bin/bash
# ncu launch list of bench.py itself (same command line, short): plain run first, then under ncu
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-sliding-window --no-cpu-baseline"
$CMD > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file gpurun_out/launches_bench.csv $CMD > gpurun_out/bench_under_ncu.json 2> gpurun_out/ncu_bench.err
cat gpurun_out/bench_plain.json | cut -c1-300
wc -l gpurun_out/launches_bench.csv
python tools/summarize_launches.py gpurun_out/launches_bench.csv 3 | head -30
