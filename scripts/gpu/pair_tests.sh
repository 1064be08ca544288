#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pair.py -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_pair.log
grep -E "^E  |FAILED|passed|failed|Error" gpurun_out/pytest_pair.log | head -30
