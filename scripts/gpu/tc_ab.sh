#!/bin/bash
# tensor-core kernel change: all GPU tests, kernel-alone timing, bench with the per-kernel breakdown
mkdir -p gpurun_out
bash scripts/gpu/tests_some.sh tests
timeout 200 python scripts/bench_tc_kernel.py 2>&1 | grep -E "DIAG|wgrad|Error|error" | head
timeout 400 python bench.py --no-extras --no-cpu-baseline --steps 60 > gpurun_out/bench_ab.json 2>/dev/null
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_ab.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "ms_per_step_median", "host_enqueue_ms_per_step")}, d["e2e"]["value"])
for k, v in list(d["kernel_breakdown"].items())[:9]:
    print(k, v if not isinstance(v, dict) else {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items()})
print(d["roofline"])
PY
