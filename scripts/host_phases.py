"""Host time of the training thread per sub-phase of net.forward / backward (perf_counter wrappers, no device sync).
usage (GPU box): python scripts/host_phases.py"""
import collections
import functools
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rslo_b200 import kernels as K  # noqa: E402
from rslo_b200.layers import encoder_engine, pose_tail, sparse3d  # noqa: E402
from rslo_b200.models import middle as middle_mod  # noqa: E402

ACC = collections.defaultdict(lambda: [0, 0.0])


def timed(name, fn):
    @functools.wraps(fn)
    def w(*a, **k):
        t = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            r = ACC[name]
            r[0] += 1
            r[1] += time.perf_counter() - t
    return w


dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
run = bench.Runner(bench.workload_config("train"), dev, 0, 1)
net = run.net
for i in range(12):
    run.step(i, False)
torch.cuda.synchronize()
net.network_forward = timed("forward.network_forward", net.network_forward)
net.loss = timed("forward.loss", net.loss)
net.create_loss = timed("forward.loss.create_loss", net.create_loss)
net._consistency_loss.forward = timed("forward.loss.create_loss.consistency", net._consistency_loss.forward)
net._loss_tail = timed("forward.loss.create_loss.loss_tail", net._loss_tail)
net.middle_feature_extractor.forward_frames = timed("forward.network_forward.middle", net.middle_feature_extractor.forward_frames)
net.odom_predictor.forward = timed("forward.network_forward.head", net.odom_predictor.forward)
encoder_engine.encode = timed("forward.network_forward.middle.engine", encoder_engine.encode)
middle_mod.build_tables_batched = timed("forward.finish_tables", middle_mod.build_tables_batched)
for name in ("nn_exact", "kth_threshold", "cov_residual", "kabsch", "spconv_tc_forward", "spconv_forward", "bn1d_seg_forward",
             "bn1d_seg_backward", "act_backward", "spconv_tc_backward_weight", "spconv_backward_weight", "spconv_backward_data",
             "table_concat", "dense_from_sites", "dense_backward", "spconv_tc_prepare"):
    setattr(K, name, timed("K." + name, getattr(K, name)))
eng_bwd = encoder_engine.SparseEncoderEngine.backward
encoder_engine.SparseEncoderEngine.backward = timed("backward.engine", eng_bwd)
N = 30
for k in run.host_phase:
    run.host_phase[k] = 0
t0 = time.perf_counter()
for i in range(N):
    run.step(20 + i, False)
host = time.perf_counter() - t0
torch.cuda.synchronize()
print(f"host {1e3 * host / N:.2f} ms/step; phases", {k: round(1e3 * v / N, 2) for k, v in run.host_phase.items() if k != "steps"})
for k, (n, t) in sorted(ACC.items()):
    print(f"  {k:52s} {n / N:6.1f} calls/step {1e3 * t / N:7.3f} ms/step  {1e6 * t / max(n, 1):7.1f} us/call")
