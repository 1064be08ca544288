#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pair.py -m gpu -q 2>&1 | tail -150 > gpurun_out/pytest_pair.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err
RSLO_BENCH_CUDA_PROFILER=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/ncu_bench.log 2>&1
grep -E "^E  |FAILED|passed|failed" gpurun_out/pytest_pair.log | head -60; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err; wc -l gpurun_out/launches.csv
