#!/bin/bash
# ncu evidence for round 2: $1 = tag (default r02).  (1) launch list of one timed step (2 pairs per GPU, the bench's
# workload), (2) --set full captures of the dominant kernels inside the same bench command.
TAG=${1:-r02}
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-extras"
if [ "$2" != "nolist" ]; then
RSLO_BENCH_CUDA_PROFILER=1 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/ncu_bench_$TAG.log 2>&1
wc -l gpurun_out/launches_$TAG.csv
fi
if [ "$2" = "listonly" ]; then exit 0; fi
full() {  # name regex count
  RSLO_BENCH_CUDA_PROFILER=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$2" -c $3 -f -o gpurun_out/prof_${TAG}_$1 $B > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log | cut -c1-200
  # gpurun brings back at most 64 MiB: keep the raw-metric page (and the hottest source lines) as CSV, drop the report
  ncu -i gpurun_out/prof_${TAG}_$1.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_$1.raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_${TAG}_$1.ncu-rep --page source --csv --print-source sass 2>/dev/null | head -4000 > gpurun_out/prof_${TAG}_$1.sass.csv
  rm -f gpurun_out/prof_${TAG}_$1.ncu-rep
}
full conv2d_tc '^k_conv2d_tc$' 24
full spconv_tc '^k_spconv_tc$' 8
full spconv_tc_wgrad 'k_spconv_tc_wgrad' 4
full conv2d_wgrad_tc 'k_conv2d_wgrad_tc' 6
full spconv_fwd '^k_spconv_fwd$' 3
full head_bn 'k_bn_act_fwd|k_bn_bwd_reduce|k_bn_bwd_apply' 6
full nn_query 'k_nn_query|k_kth_threshold' 4
ls -la gpurun_out/*.csv
