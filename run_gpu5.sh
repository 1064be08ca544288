#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "tensor_core or wgrad" 2>&1 | tail -40 > gpurun_out/pytest_tc.log
tail -40 gpurun_out/pytest_tc.log
nvidia-smi --query-gpu=name,memory.used --format=csv
