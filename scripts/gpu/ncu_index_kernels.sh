#!/bin/bash
# ncu --set full of the indexing kernels (voxeliser, rulebook, NN) inside one timed bench step
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --pairs-per-gpu 1 --no-cpu-baseline --no-profile"
N="ncu --set full --clock-control none --import-source on --profile-from-start off -f"
RSLO_BENCH_CUDA_PROFILER=1 timeout 600 $N -k regex:'^k_vox_' -c 8 -o gpurun_out/prof_r01_voxelize $B > gpurun_out/ncu_vox.log 2>&1; tail -1 gpurun_out/ncu_vox.log
RSLO_BENCH_CUDA_PROFILER=1 timeout 600 $N -k regex:'^k_(subm_table|strided_table|strided_mark|site_mark)$' -c 8 -o gpurun_out/prof_r01_rulebook $B > gpurun_out/ncu_rb.log 2>&1; tail -1 gpurun_out/ncu_rb.log
RSLO_BENCH_CUDA_PROFILER=1 timeout 600 $N -k regex:'^k_(nn_query|nn_fill|cov_fwd|cov_bwd|kabsch_accum)$' -c 8 -o gpurun_out/prof_r01_loss_kernels $B > gpurun_out/ncu_nn.log 2>&1; tail -1 gpurun_out/ncu_nn.log
ls -la gpurun_out/prof_r01_*.ncu-rep
