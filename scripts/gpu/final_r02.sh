#!/bin/bash
# round-2 final state: GPU tests, smoke, default bench (all extras), launch list + --set full of the two dominant kernels
TAG=${1:-r02g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_all.log; tail -1 gpurun_out/pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
timeout 900 python bench.py > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; cut -c1-600 gpurun_out/bench_train.json; tail -2 gpurun_out/bench_train.err | cut -c1-200
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-extras"
RSLO_BENCH_CUDA_PROFILER=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/ncu_bench_$TAG.log 2>&1
wc -l gpurun_out/launches_$TAG.csv
full() {  # name regex count
  RSLO_BENCH_CUDA_PROFILER=1 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$2" -c $3 -f -o gpurun_out/prof_${TAG}_$1 $B > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log | cut -c1-200
  ncu -i gpurun_out/prof_${TAG}_$1.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_$1.raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_${TAG}_$1.ncu-rep --page source --csv --print-source sass 2>/dev/null | head -4000 > gpurun_out/prof_${TAG}_$1.sass.csv
  rm -f gpurun_out/prof_${TAG}_$1.ncu-rep
}
full conv2d_tc '^k_conv2d_tc$' 12
full spconv_tc '^k_spconv_tc$' 8
