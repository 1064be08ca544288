"""2-rank NCCL check (launched by tests/test_gpu_multi.py through torchrun): after one training step on DIFFERENT
samples per rank, FlatGradAllReducer.all_reduce() leaves on every rank the mean of the two ranks' local gradients
(early head bucket overlapped with the encoder's backward included), and sync_buffers() equalises the BatchNorm
statistics."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rslo_b200  # noqa: E402
from rslo_b200.data import synthetic  # noqa: E402
from rslo_b200.utils.distributed import FlatGradAllReducer, init_from_env  # noqa: E402
from rslo_b200.utils.weights import deterministic_fill  # noqa: E402

rank, local, world = init_from_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
net, _ = rslo_b200.build_network(testing=False, seed=7)
deterministic_fill(net, 11)
net = net.to(dev).train()
net.global_step.fill_(2000)
net._step_host = None
red = FlatGradAllReducer(net, early_module=net.odom_predictor)
red.broadcast_params()
a, b, _ = synthetic.make_pair(50 + rank, n_beams=16, n_az=600)
ex = {"points": [torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)], "host_outputs": False}

# local gradients first (no reducer hook): plain backward, cloned
net.on_head_backward_done = None
red.zero_()
net(ex)["loss"].sum().backward()
local_g = {k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None}
# now the real cycle, with the early bucket launched from the autograd hook
net.on_head_backward_done = red._reduce_early
for _ in range(2):                      # second round = steady state (graph replays, hooks re-armed)
    red.zero_()
    net(ex)["loss"].sum().backward()
    assert red._early_done, "the head-backward-done hook did not fire"
    red.all_reduce()
torch.cuda.synchronize()
worst, worst_key = 0.0, None
for k, p in net.named_parameters():
    if k not in local_g:
        continue
    gs = [torch.empty_like(local_g[k]) for _ in range(world)]
    dist.all_gather(gs, local_g[k])
    mean = sum(g.double() for g in gs) / world
    err = float((p.grad.double() - mean).abs().max() / mean.abs().max().clamp_min(1e-30))
    # parameters whose true gradient is zero (biases of the softmax logits, biases ahead of a batch-statistics
    # BatchNorm) carry 1e-11-level rounding noise only: nothing to compare
    if float(mean.abs().max()) > 1e-7 and err > worst:
        worst, worst_key = err, k
# BN statistics: per rank during training, equal after sync_buffers
rm = net.odom_predictor.blocks[0][0].bn1.running_mean
both = [torch.empty_like(rm) for _ in range(world)]
dist.all_gather(both, rm)
differ = float((both[0] - both[1]).abs().max())
red.sync_buffers(net)
dist.all_gather(both, rm)
same = float((both[0] - both[1]).abs().max())
# fused optimizer step on the summed gradients (1/world folded into the kernel): every rank ends with the same weights
from rslo_b200.torchplus.train import FusedAdamClip  # noqa: E402
opt = FusedAdamClip(red, lr=1e-3, wd=1e-5, max_norm=10.0)
red.zero_()
net(ex)["loss"].sum().backward()
red.all_reduce(average=False)
assert red.pending_average
norm = opt.clip_grad_norm_()
opt.step()
w = net.odom_predictor.blocks[0][0].conv1.conv1.weight.detach()
ws_ = [torch.empty_like(w) for _ in range(world)]
dist.all_gather(ws_, w.contiguous())
norms = [torch.empty_like(norm) for _ in range(world)]
dist.all_gather(norms, norm)
opt_same = bool(torch.equal(ws_[0], ws_[1])) and bool(torch.equal(norms[0], norms[1]))
if rank == 0:
    assert opt_same, "optimizer step diverged across ranks"
    # graph replays are bitwise repeatable except for the double-precision atomics of the BN statistics
    print(f"DDP_CHECK worst_rel_err={worst:.3e} ({worst_key}) bn_differ_before={differ:.3e} bn_differ_after={same:.3e}", flush=True)
    assert worst < 5e-3, worst
    assert differ > 0 and same == 0.0
dist.barrier()
dist.destroy_process_group()
