#!/bin/bash
# first-light + timing of the dense TMA/tcgen05 convolutions; each mode in its own process (a trap kills the context)
mkdir -p gpurun_out
for m in fwd dgrad wgrad; do
  timeout 300 python scripts/conv2d_check.py $m --time > gpurun_out/conv2d_$m.log 2>&1
  echo "== $m rc=$?"; tail -n 70 gpurun_out/conv2d_$m.log | cut -c1-400
done
