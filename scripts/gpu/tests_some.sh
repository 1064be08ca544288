#!/bin/bash
# $@ = pytest selectors; full short tracebacks kept in gpurun_out/pytest_some.log
mkdir -p gpurun_out
timeout 1500 python -m pytest "$@" -m gpu -q --tb=short -W ignore 2>&1 > gpurun_out/pytest_some.log; grep -E "^(FAILED|ERROR|E  )|passed|failed" gpurun_out/pytest_some.log | cut -c1-300 | tail -60
