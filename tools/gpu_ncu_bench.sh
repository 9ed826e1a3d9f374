#!/bin/bash
# usage: gpu_ncu_bench.sh <tag>   ncu launch list of bench.py (first ~2 training steps: same kernels as the replayed graph)
mkdir -p gpurun_out
TAG=${1:-r02}
CMD="python bench.py --steps 2 --warmup 1 --no-sliding-window --no-cpu-baseline --no-library-gpu --no-hybrid --no-class-probe"
$CMD > gpurun_out/bench_plain_$TAG.json 2> gpurun_out/bench_plain_$TAG.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3800 --csv --log-file gpurun_out/launches_bench_$TAG.csv $CMD > gpurun_out/bench_under_ncu.json 2> gpurun_out/ncu_bench.err
cut -c1-200 gpurun_out/bench_plain_$TAG.json
python tools/step_from_launches.py gpurun_out/launches_bench_$TAG.csv | tee gpurun_out/launches_bench_${TAG}_summary.txt
