#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu/tests_some.sh tests
for v in 1 0; do
  RSLO_PDL=$v timeout 300 python bench.py --no-extras --no-cpu-baseline --no-profile --steps 60 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('RSLO_PDL=$v', {k: round(d[k],3) for k in ('value','ms_per_step','ms_per_step_median','host_enqueue_ms_per_step')}, round(d['e2e']['value'],2))"
done
