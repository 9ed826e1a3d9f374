#!/bin/bash
# usage: gpu_prof.sh <tag> [batch]
mkdir -p gpurun_out
B=${2:-4}
python tools/profile_forward.py $B > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$1.csv python tools/profile_forward.py $B > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log; wc -l gpurun_out/launches_$1.csv
