#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err
cut -c1-1500 gpurun_out/bench_n2.log; grep -iE "error|Traceback" gpurun_out/bench_n2.err | head -5 | cut -c1-300
