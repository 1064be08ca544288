#!/bin/bash
mkdir -p gpurun_out
RSLO_BENCH_CUDA_PROFILER=1 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'^k_spconv_tc$' -c 6 -f -o gpurun_out/prof_r01_spconv_tc python bench.py --steps 1 --warmup 3 --pairs-per-gpu 1 --no-cpu-baseline --no-profile > gpurun_out/ncu_tc.log 2>&1
tail -2 gpurun_out/ncu_tc.log; ls -la gpurun_out/prof_r01_*.ncu-rep
