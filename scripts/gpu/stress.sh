#!/bin/bash
mkdir -p gpurun_out

timeout 600 python bench.py --workload stress --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_stress.log 2> gpurun_out/bench_stress.err
cut -c1-1500 gpurun_out/bench_stress.log; grep -E "Error|error|Traceback" -A5 gpurun_out/bench_stress.err | head -20 | cut -c1-300
