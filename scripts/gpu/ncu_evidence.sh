#!/bin/bash
mkdir -p gpurun_out
# (1) every launch of one timed step (1 pair per GPU to bound the replay time)
RSLO_BENCH_CUDA_PROFILER=1 timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 3 --pairs-per-gpu 1 --no-cpu-baseline --no-profile > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/launches_r01.csv
# (2) full capture of the dominant own kernels inside the same command
RSLO_BENCH_CUDA_PROFILER=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'^k_spconv_tc$' -c 6 -f -o gpurun_out/prof_r01_spconv_tc python bench.py --steps 1 --warmup 3 --pairs-per-gpu 1 --no-cpu-baseline --no-profile > gpurun_out/ncu_tc.log 2>&1
tail -2 gpurun_out/ncu_tc.log
RSLO_BENCH_CUDA_PROFILER=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_spconv_tc_wgrad -c 3 -f -o gpurun_out/prof_r01_spconv_tc_wgrad python bench.py --steps 1 --warmup 3 --pairs-per-gpu 1 --no-cpu-baseline --no-profile > gpurun_out/ncu_wg.log 2>&1
tail -2 gpurun_out/ncu_wg.log
ls -la gpurun_out/*.ncu-rep
