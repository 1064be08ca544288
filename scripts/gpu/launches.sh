#!/bin/bash
# every kernel launch of one timed step (ncu launch list); $1 = tag, $2 = pairs per gpu
TAG=${1:-r02}; PPG=${2:-2}
mkdir -p gpurun_out
RSLO_BENCH_CUDA_PROFILER=1 timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --pairs-per-gpu $PPG --no-cpu-baseline --no-profile > gpurun_out/ncu_bench_$TAG.log 2>&1
wc -l gpurun_out/launches_$TAG.csv; tail -3 gpurun_out/ncu_bench_$TAG.log | cut -c1-300
