#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pair.py -m gpu -q -x 2>&1 | tail -80 > gpurun_out/pytest_pair.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err
RSLO_CUDA_GRAPHS=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/bench_nograph.log 2> gpurun_out/bench_nograph.err
grep -E "^E  |FAILED|passed|failed|Error" gpurun_out/pytest_pair.log | head -40; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err; cat gpurun_out/bench_nograph.log
