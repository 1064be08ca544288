"""The odometry head's trunk on the repo's own kernels (layers/head_tc.py over csrc/conv2d_tc.cu + csrc/head_ops.cu)
against the same module tree evaluated by torch in FLOAT64 (`_trunk_torch`, the reference's layer sequence
`rslo/models/odom_pred.py:152-260`): outputs, input gradients, every parameter gradient, BatchNorm running
statistics; training mode with per-sample statistics groups, frozen-BN mode and eval mode.
Tolerance: 2e-5 of the tensor's max magnitude (split-TF32 products are FP32-level; measured ~1e-6)."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def head(cuda):
    import rslo_b200
    from rslo_b200.utils.weights import deterministic_fill
    net, _ = rslo_b200.build_network(testing=False, seed=7)
    deterministic_fill(net, 11)
    h = net.odom_predictor.cuda()
    # non-trivial BN affine parameters / running statistics
    g = torch.Generator(device="cpu").manual_seed(3)
    for m in h.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data = (0.5 + torch.rand(m.num_features, generator=g)).cuda()
            m.bias.data = (0.2 * torch.randn(m.num_features, generator=g)).cuda()
            m.running_mean.data = (0.1 * torch.randn(m.num_features, generator=g)).cuda()
            m.running_var.data = (0.5 + torch.rand(m.num_features, generator=g)).cuda()
    return h


def _inputs(S, seed, H=96, W=176, C=128):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x1 = torch.randn(S, C, H, W, generator=g)
    x2 = torch.randn(S, C, H, W, generator=g)
    occ1 = (torch.rand(S, 1, H, W, generator=g) < 0.35).float()
    occ2 = (torch.rand(S, 1, H, W, generator=g) < 0.35).float()
    return (x1 * occ1).cuda(), (x2 * occ2).cuda()


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _ref_trunk(head, x1, x2, groups):
    """float64 torch evaluation, one statistics group (sample) at a time like the reference"""
    stash = {k: head.__dict__.pop(k) for k in ("_trunk_engine", "_graphed") if k in head.__dict__}
    h64 = copy.deepcopy(head).double()
    head.__dict__.update(stash)
    outs, masks = None, []
    xs = []
    for s in range(groups):
        n = x1.shape[0] // groups
        a = x1[s * n:(s + 1) * n].double().detach().requires_grad_(True)
        b = x2[s * n:(s + 1) * n].double().detach().requires_grad_(True)
        xs.append((a, b))
        tq, tl, rl, py, mask = h64._trunk_torch(a, b)
        o = [tq, tl, rl] + py
        outs = [[v] for v in o] if outs is None else [p + [v] for p, v in zip(outs, o)]
        masks.append(mask)
    return h64, [torch.cat(v) for v in outs], torch.cat(masks), xs


def _own(head, x1, x2, ipg):
    x1 = x1.detach().requires_grad_(True)
    x2 = x2.detach().requires_grad_(True)
    tq, tl, rl, py, mask = head._trunk_own(x1, x2, ipg)
    return [tq, tl, rl] + py, mask, (x1, x2)


def _loss(outs, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    tot = 0
    for o in outs:
        w = torch.randn(o.shape, generator=g).to(o.device, o.dtype)
        tot = tot + (o * w).sum()
    return tot


@pytest.mark.parametrize("mode,S", [("train", 2), ("train", 1), ("frozen", 2), ("eval", 1)])
def test_trunk_matches_float64(head, mode, S):
    head.train(mode != "eval")
    if mode == "frozen":                      # freeze_bn: every BN layer in eval mode while the net trains
        for m in head.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.eval()
    x1, x2 = _inputs(S, 5 + S)
    state0 = copy.deepcopy(head.state_dict())
    h64, ref, ref_mask, ref_x = _ref_trunk(head, x1, x2, S)
    head.load_state_dict(state0)
    outs, mask, (a, b) = _own(head, x1, x2, 1)
    assert torch.equal(mask, ref_mask.float())
    names = ["tq_map", "t_logit", "r_logit", "py0", "py1"]
    for n, o, r in zip(names, outs, ref):
        assert o.shape == r.shape, n
        assert _rel(o, r) < 2e-5, (n, _rel(o, r))
    # running statistics after one forward
    sd, sd64 = head.state_dict(), h64.state_dict()
    for k in sd:
        if "running_" in k or "num_batches" in k:
            if sd[k].dtype == torch.long:
                assert int(sd[k]) == int(sd64[k]), k
            else:
                assert _rel(sd[k], sd64[k]) < 1e-5, k
    if mode == "eval":
        return
    _loss(outs, 1).backward()
    _loss(ref, 1).backward()
    gx1 = torch.cat([p[0].grad for p in ref_x])
    gx2 = torch.cat([p[1].grad for p in ref_x])
    assert _rel(a.grad, gx1) < 2e-5, _rel(a.grad, gx1)
    assert _rel(b.grad, gx2) < 2e-5, _rel(b.grad, gx2)
    p64 = dict(h64.named_parameters())
    worst = ("", 0.0)
    for k, p in head.named_parameters():
        r = p64[k].grad
        if r is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        assert p.grad is not None, k
        scale = float(r.abs().max())
        if k.endswith(".bias") and scale < 1e-6 * max(1.0, float(p64[k[:-5] + ".weight"].grad.abs().max())):
            continue                           # bias ahead of a batch-statistics BN: the true gradient is 0 (noise)
        e = _rel(p.grad, r)
        if e > worst[1]:
            worst = (k, e)
    assert worst[1] < 5e-5, worst
    head.zero_grad()


def test_trunk_graph_replay_matches_eager(head):
    """forward/backward through the CUDA-graph path (make_graphed_callables) == eager engine calls"""
    head.train()
    x1, x2 = _inputs(1, 9)
    st = copy.deepcopy(head.state_dict())
    xs = [x1.detach().requires_grad_(True), x2.detach().requires_grad_(True)]
    head.use_cuda_graph = False
    d0 = head(xs)
    (d0["translation_preds"][0].sum() + d0["rotation_preds"][0].sum() + d0["pyramid_motion"][0][0].sum()).backward()
    g_ref = {k: p.grad.clone() for k, p in head.named_parameters() if p.grad is not None}
    gx_ref = xs[0].grad.clone()
    head.zero_grad()
    head.load_state_dict(st)
    head.use_cuda_graph = True
    for _ in range(2):                       # second call = pure replay
        ys = [x1.detach().requires_grad_(True), x2.detach().requires_grad_(True)]
        head.load_state_dict(st)
        head.zero_grad()
        d1 = head(ys)
        (d1["translation_preds"][0].sum() + d1["rotation_preds"][0].sum() + d1["pyramid_motion"][0][0].sum()).backward()
    assert _rel(d1["translation_preds"][0], d0["translation_preds"][0]) < 1e-6
    assert _rel(ys[0].grad, gx_ref) < 1e-5
    for k, p in head.named_parameters():
        if k in g_ref and float(g_ref[k].abs().max()) > 0:
            assert _rel(p.grad, g_ref[k]) < 1e-4, k
    head.zero_grad()


def test_weight_update_through_data_is_seen(head):
    """in-place writes through .data (optimizer wrappers, broadcast) must not leave stale weight images"""
    head.eval()
    x1, x2 = _inputs(1, 4)
    with torch.no_grad():
        o0 = head._trunk_own(x1, x2, 1)[0].clone()
        w = head.tq_map_conv[0].weight
        w.data.mul_(1.5)
        o1 = head._trunk_own(x1, x2, 1)[0].clone()
        w.data.div_(1.5)
    assert float((o1 - o0).abs().max()) > 1e-6
