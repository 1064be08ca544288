#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/profile_step.py --pairs 2 --steps 3 > gpurun_out/profile_step.txt 2>&1; head -5 gpurun_out/profile_step.txt; grep -n "profiler:" gpurun_out/profile_step.txt
for w in fwd dgrad wgrad; do timeout 300 python scripts/conv2d_check.py $w --time > gpurun_out/conv2d_$w.txt 2>&1; done; tail -2 gpurun_out/conv2d_fwd.txt
