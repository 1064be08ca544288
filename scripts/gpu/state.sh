#!/bin/bash
# state check: GPU tests, smoke, default bench, launch list of one step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_all.log; tail -3 gpurun_out/pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
timeout 900 python bench.py > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; cut -c1-3000 gpurun_out/bench_train.json; tail -3 gpurun_out/bench_train.err
bash scripts/gpu/ncu_r02.sh r02f listonly
