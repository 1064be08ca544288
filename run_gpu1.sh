#!/bin/bash
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_pair.py 2>&1 | tail -60 > gpurun_out/pytest_kernels.log
timeout 900 python -m pytest tests/test_gpu_pair.py -m gpu -q 2>&1 | tail -120 > gpurun_out/pytest_pair.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
tail -5 gpurun_out/pytest_kernels.log; tail -30 gpurun_out/pytest_pair.log; cat gpurun_out/bench.log; tail -20 gpurun_out/bench.err
