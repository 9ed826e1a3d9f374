#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run8.log
: > $L
timeout 1800 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -15 >> $L
timeout 600 python tools/time_forward.py >> $L 2>&1
timeout 600 python tools/profile_sections.py 4 2>&1 | head -24 >> $L
tail -60 $L
