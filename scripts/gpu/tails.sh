#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tail.py -q -x 2>&1 | grep -v Warning | tail -30
timeout 900 python -m pytest tests/test_gpu_pair.py tests/test_gpu_head.py -q -x 2>&1 | grep -v Warning | tail -15
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_tail.json 2> gpurun_out/bench_tail.err; tail -c 300 gpurun_out/bench_tail.err; cut -c1-200 gpurun_out/bench_tail.json
