"""Where does the training thread's HOST time go?  cProfile over bench.py's train step (no device sync inside).
usage (GPU box): python scripts/host_profile.py [--steps 20] > gpurun_out/host_profile.txt"""
import argparse
import cProfile
import io
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("RSLO_BENCH_PREFETCH_THREAD", "0")
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=20)
args = ap.parse_args()
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
run = bench.Runner(bench.workload_config("train"), dev, 0, 1)
for i in range(12):
    run.step(i, False)
torch.cuda.synchronize()
t0 = time.time()
for i in range(args.steps):
    run.step(12 + i, False)
t_host = time.time() - t0
torch.cuda.synchronize()
print(f"host enqueue {1e3 * t_host / args.steps:.2f} ms/step, wall {1e3 * (time.time() - t0) / args.steps:.2f} ms/step")
pr = cProfile.Profile()
pr.enable()
for i in range(args.steps):
    run.step(40 + i, False)
pr.disable()
torch.cuda.synchronize()
for key in ("cumulative", "tottime"):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats(key).print_stats(45)
    txt = s.getvalue().replace(ROOT + "/", "")
    print(txt[:9000])
