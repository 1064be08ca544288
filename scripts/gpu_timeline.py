"""GPU timeline of bench.py's train step (torch.profiler / kineto; no nsys in the image): per stream busy time, the
idle gaps of the whole device inside a step and which kernels border them.
usage (GPU box): python scripts/gpu_timeline.py [--steps 3] > gpurun_out/timeline.txt"""
import argparse
import json
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--workload", default="train")
args = ap.parse_args()

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
run = bench.Runner(bench.workload_config(args.workload), dev, 0, 1)
for i in range(12):
    run.step(i, False)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(args.steps):
        run.flush.zero_()
        run.step(12 + i, False)
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
ev = json.load(open(path))["traceEvents"]
gpu = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
gpu.sort(key=lambda e: e["ts"])
t0, t1 = gpu[0]["ts"], max(e["ts"] + e["dur"] for e in gpu)
print(f"{len(gpu)} GPU activities over {(t1 - t0) / 1e3:.2f} ms = {(t1 - t0) / 1e3 / args.steps:.2f} ms/step")
streams = {}
for e in gpu:
    streams.setdefault(e["args"].get("stream"), []).append(e)
for s, es in sorted(streams.items(), key=lambda kv: -sum(e["dur"] for e in kv[1])):
    print(f"  stream {s}: {len(es) / args.steps:.0f} activities/step, busy {sum(e['dur'] for e in es) / 1e3 / args.steps:.2f} ms/step")
# union of busy intervals over all streams -> device idle gaps
busy_end = gpu[0]["ts"]
gaps = []
prev = gpu[0]
for e in gpu:
    if e["ts"] > busy_end:
        gaps.append((e["ts"] - busy_end, prev["name"][:60], e["name"][:60], e["ts"] - t0))
    if e["ts"] + e["dur"] > busy_end:
        busy_end = e["ts"] + e["dur"]
        prev = e
idle = sum(g[0] for g in gaps)
print(f"device idle {idle / 1e3 / args.steps:.2f} ms/step in {len(gaps) / args.steps:.0f} gaps/step; gaps >= 10 us: "
      f"{sum(g[0] for g in gaps if g[0] >= 10) / 1e3 / args.steps:.2f} ms/step")
hist = {}
for d, a, b, _ in gaps:
    k = (a, b)
    h = hist.setdefault(k, [0, 0.0])
    h[0] += 1
    h[1] += d
print("largest idle contributors (kernel before -> kernel after): count/step, idle us/step")
for (a, b), (n, d) in sorted(hist.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"  {n / args.steps:6.1f} {d / args.steps:8.1f}  {a}  ->  {b}")
# host: time of the python step() calls
cpu = [e for e in ev if e.get("cat") in ("cpu_op", "python_function", "user_annotation", "cuda_runtime") and "dur" in e]
launch = [e for e in cpu if e.get("cat") == "cuda_runtime"]
print(f"cuda runtime calls/step: {len(launch) / args.steps:.0f}, host time in them {sum(e['dur'] for e in launch) / 1e3 / args.steps:.2f} ms/step")

# which torch ops launch the small kernels: aten ops by (name, input shapes), calls per step
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof2:
    for i in range(args.steps):
        run.flush.zero_()
        run.step(40 + i, False)
    torch.cuda.synchronize()
rows = []
for e in prof2.key_averages(group_by_input_shape=True):
    if e.key.startswith("aten::") and e.self_device_time_total > 0:
        rows.append((e.count / args.steps, e.self_device_time_total / args.steps, e.key, str(e.input_shapes)[:110]))
rows.sort(key=lambda r: -r[0])
print("aten ops that launch kernels: calls/step, device us/step, op, input shapes")
for r in rows[:70]:
    print(f"  {r[0]:6.1f} {r[1]:8.1f}  {r[2]:34s} {r[3]}")
print(f"  total aten kernel-launching calls/step: {sum(r[0] for r in rows):.0f}, device time {sum(r[1] for r in rows) / 1e3:.2f} ms/step")
