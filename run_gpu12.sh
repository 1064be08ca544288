#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s -k "tensor_core" 2>&1 | tail -60 > gpurun_out/pytest_tc.log
grep -E "tc max|^E  |FAILED|passed|failed|Error" gpurun_out/pytest_tc.log | head -40
timeout 900 python -m pytest tests/test_gpu_pair.py -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_pair_tc.log
grep -E "^E  |FAILED|passed|failed|Error" gpurun_out/pytest_pair_tc.log | head -40
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err
cut -c1-400 gpurun_out/bench.log
timeout 600 python scripts/profile_step.py --pairs 2 > gpurun_out/profile_step.txt 2>&1
grep -E "wall ms|profiler:|Self CUDA time total" gpurun_out/profile_step.txt
grep -E "k_spconv_tc" gpurun_out/profile_step.txt | head -5 | cut -c1-100,150-250
