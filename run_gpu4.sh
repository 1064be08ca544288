#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/profile_step.py --pairs 2 > gpurun_out/profile_step.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "nn_bit_exact" 2>&1 | tail -5
head -12 gpurun_out/profile_step.txt
