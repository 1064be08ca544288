#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s -k "tensor_core_wgrad" 2>&1 | tail -40 > gpurun_out/pytest_wg.log
grep -E "tc wgrad|^E  |FAILED|passed|failed|Error" gpurun_out/pytest_wg.log | head -40
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_all.log
grep -E "^E  |FAILED|passed|failed|Error" gpurun_out/pytest_all.log | head -40
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err
cut -c1-400 gpurun_out/bench.log; tail -3 gpurun_out/bench.err | cut -c1-300
timeout 600 python scripts/profile_step.py --pairs 2 > gpurun_out/profile_step.txt 2>&1
grep -E "wall ms|profiler:|Self CUDA time total" gpurun_out/profile_step.txt
grep -E "^(void |rslo|cutlass|std::)" gpurun_out/profile_step.txt | head -14 | cut -c1-100,150-250
