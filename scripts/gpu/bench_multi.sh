#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
cut -c1-700 gpurun_out/bench_n$N.log; grep -iE "error|Traceback" gpurun_out/bench_n$N.err | head -5 | cut -c1-300
