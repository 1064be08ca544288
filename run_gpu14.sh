#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s -x -k "tensor_core_wgrad or kth" 2>&1 | tail -40 > gpurun_out/pytest_wg.log
grep -E "tc wgrad|^E  |FAILED|passed|failed|Error" gpurun_out/pytest_wg.log | head -40
