#!/bin/bash
# one iteration: selected tests, then the quick bench line (no CPU baseline / extras)
mkdir -p gpurun_out
bash scripts/gpu/tests_some.sh ${TESTS:-tests}
QB="--steps 40 --warmup 10 --no-cpu-baseline --no-extras"
for g in ${GROUPS_TO_TRY:-1}; do
  RSLO_CONV_DRAIN_GROUP=$g timeout 600 python bench.py $QB > gpurun_out/bench_quick_g$g.json 2> gpurun_out/bench_quick_g$g.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_quick_g$g.json"))
print("group $g:", round(d["value"],2), "pairs/s; ms/step mean", round(d["ms_per_step"],3), "median", round(d["ms_per_step_median"],3), "p90/max", [round(v,2) for v in d.get("ms_per_step_p90_max",[])], "; e2e", round(d["e2e"]["value"],2), "; host enqueue ms", round(d.get("host_enqueue_ms_per_step",0),2), {k: round(v,2) for k,v in d.get("host_phase_ms_per_step",{}).items()}, "| e2e phases", {k: round(v,2) for k,v in d["e2e"].get("host_phase_ms_per_step",{}).items()}, d["e2e"].get("slots"))
kb=d.get("kernel_breakdown",{})
for k,v in sorted(((k,v) for k,v in kb.items() if isinstance(v,dict)), key=lambda kv:-kv[1]["ms_per_step"])[:8]:
    print("   %-26s %6.1f %7.3f" % (k, v["calls_per_step"], v["ms_per_step"]))
PY
done
