"""Where does a training step's time go?  torch.profiler over a few steps of bench.py's workload:
GPU-busy time per kernel group and host time per op; also wall time per stage with synchronisation.
Usage (GPU box): python scripts/profile_step.py [--pairs 1] > gpurun_out/profile_step.txt"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import rslo_b200  # noqa: E402
from oracle import net as onet  # noqa: E402
from rslo_b200 import kernels as K  # noqa: E402
from rslo_b200.data import synthetic  # noqa: E402
from rslo_b200.utils.distributed import FlatGradAllReducer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=2)
ap.add_argument("--steps", type=int, default=3)
args = ap.parse_args()

dev = torch.device("cuda", 0)
net, vg = rslo_b200.build_network(testing=False, seed=7)
onet.fill_weights(net, 11)
net = net.to(dev)
net.global_step.fill_(2000)
net._step_host = None
net.train()
red = FlatGradAllReducer(net)
pairs = [synthetic.make_pair(s)[:2] for s in range(args.pairs)]
pairs = [(torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)) for a, b in pairs]


def step():
    red.zero_()
    pts = []
    for a, b in pairs:
        pts += [a, b]
    ret = net({"points": pts, "n_samples": len(pairs), "host_outputs": False})
    ret["loss"].sum().backward()


for _ in range(4):
    step()
torch.cuda.synchronize()
t0 = time.time()
for _ in range(args.steps):
    step()
t_host = time.time() - t0
torch.cuda.synchronize()
print(f"wall ms/step {1e3 * (time.time() - t0) / args.steps:.2f}  ({args.pairs} pairs); host time to ENQUEUE a step "
      f"{1e3 * t_host / args.steps:.2f} ms (if ~= wall, the step is launch-bound)")

# stage walls (each stage synchronised): forward pieces of ONE pair
a, b = pairs[0]
def sync_time(fn, n=5):
    torch.cuda.synchronize(); t = time.time()
    for _ in range(n):
        out = fn()
    torch.cuda.synchronize()
    return 1e3 * (time.time() - t) / n, out

ms, vox = sync_time(lambda: [net._voxelize_on_device(p) for p in (a, b)])
print(f"stage voxelize x2 frames: {ms:.2f} ms")
def enc():
    return net.middle_feature_extractor.forward_frames([v[0] for v in vox], [v[1] for v in vox], 1, [v[3] for v in vox], [v[4] for v in vox])
ms, encs = sync_time(enc)
print(f"stage sparse encoder fwd x2 frames: {ms:.2f} ms")
ms, head = sync_time(lambda: net.odom_predictor(list(encs[0])))
print(f"stage head fwd: {ms:.2f} ms")

from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
ka = prof.key_averages()
tot_cuda = sum(e.self_device_time_total for e in ka) / 1e3 / args.steps
tot_cpu = sum(e.self_cpu_time_total for e in ka) / 1e3 / args.steps
print(f"profiler: GPU-busy {tot_cuda:.2f} ms/step, host self-time {tot_cpu:.2f} ms/step")
print(ka.table(sort_by="self_device_time_total", row_limit=45, max_name_column_width=70))
print(ka.table(sort_by="self_cpu_time_total", row_limit=30, max_name_column_width=70))
