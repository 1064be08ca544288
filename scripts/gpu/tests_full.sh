#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -v Warning | tail -60 > gpurun_out/pytest_full.log; tail -25 gpurun_out/pytest_full.log | cut -c1-400
