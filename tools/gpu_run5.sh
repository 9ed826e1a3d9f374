#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/time_forward.py > gpurun_out/run5_time.log 2>&1
tail -8 gpurun_out/run5_time.log
python tools/profile_forward.py 4 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_fwd_b4.csv python tools/profile_forward.py 4 > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log; wc -l gpurun_out/launches_fwd_b4.csv
