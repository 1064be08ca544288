#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu/tests_some.sh tests/test_gpu_kernels.py tests/test_gpu_parity_holes.py tests/test_gpu_pair.py
timeout 200 python scripts/bench_tc_kernel.py 2>&1 | grep -E "DIAG|wgrad|Error|error" | head
timeout 300 python bench.py --no-extras --no-cpu-baseline --no-profile --steps 60 2>/dev/null | cut -c1-330
