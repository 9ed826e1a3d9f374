#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_r01_a.json 2> gpurun_out/bench_err.log; tail -5 gpurun_out/bench_err.log; cat gpurun_out/bench_r01_a.json
