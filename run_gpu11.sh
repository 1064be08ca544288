#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spconv_tc -c 2 -f -o gpurun_out/prof_spconv_tc python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "tensor_core and 19594" > gpurun_out/ncu_tc.log 2>&1
tail -3 gpurun_out/ncu_tc.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spconv_wgrad -c 1 -f -o gpurun_out/prof_spconv_wgrad python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "forward_and_wgrad and 64-64-27" > gpurun_out/ncu_wgrad.log 2>&1
tail -3 gpurun_out/ncu_wgrad.log
ls -la gpurun_out/*.ncu-rep
