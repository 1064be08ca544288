#!/bin/bash
# head trunk on own kernels: kernel tests, trunk-vs-float64 tests, end-to-end goldens, short bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_head.py -q -x 2>&1 | tail -30
timeout 900 python -m pytest tests/test_gpu_pair.py -q -x 2>&1 | tail -15
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_head.json 2> gpurun_out/bench_head.err; tail -c 600 gpurun_out/bench_head.err; cut -c1-900 gpurun_out/bench_head.json
timeout 300 python scripts/profile_step.py --pairs 2 > gpurun_out/profile_step_r02b.txt 2>&1; grep "wall\|stage\|profiler:" gpurun_out/profile_step_r02b.txt
