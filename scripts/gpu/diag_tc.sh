#!/bin/bash
# timing-only diagnostics of the tensor-core sparse conv: which stage of the pipeline paces it
for d in 0 1 2 3 4 5; do
  touch rslo_b200/csrc/spconv_tc.cu
  make -C rslo_b200/csrc EXTRA=-DTC_DIAG=$d > /dev/null 2>&1
  TC_DIAG=$d timeout 120 python scripts/bench_tc_kernel.py 2>&1 | grep DIAG
done
touch rslo_b200/csrc/spconv_tc.cu; make -C rslo_b200/csrc > /dev/null 2>&1
