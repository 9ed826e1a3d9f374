#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run9.log
: > $L
timeout 1800 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -15 >> $L
timeout 300 python __graft_entry__.py smoke >> $L 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_train_a.json 2>> $L
cat gpurun_out/bench_train_a.json >> $L
tail -40 $L
