#!/bin/bash
# usage: gpu_prof_train.sh <tag> [batch]
mkdir -p gpurun_out
B=${2:-2}
python tools/profile_train.py $B > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$1.csv python tools/profile_train.py $B > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log; wc -l gpurun_out/launches_$1.csv
python tools/summarize_launches.py gpurun_out/launches_$1.csv 40 | head -120
