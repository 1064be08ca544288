#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err
cut -c1-900 gpurun_out/bench_n2.log; tail -5 gpurun_out/bench_n2.err | cut -c1-300
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 --cpu-budget-s 60 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
cut -c1-900 gpurun_out/bench_ref.log; tail -3 gpurun_out/bench_ref.err
timeout 600 python bench.py --workload eval --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_eval.log 2> gpurun_out/bench_eval.err
cut -c1-700 gpurun_out/bench_eval.log; tail -3 gpurun_out/bench_eval.err | cut -c1-300
