#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "conv2d_tensor_core or tensor_core_forward" 2>&1 | tail -30 > gpurun_out/pytest_head.log
grep -E "^E  |FAILED|passed|failed|Error" gpurun_out/pytest_head.log | head -30
