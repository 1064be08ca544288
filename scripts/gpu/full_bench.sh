#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; cut -c1-400 gpurun_out/bench_train.json; tail -2 gpurun_out/bench_train.err | cut -c1-300
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json
