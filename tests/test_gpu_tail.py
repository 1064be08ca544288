"""The fused tail kernels (csrc/pose_tail.cu via layers/pose_tail.py) against the same steps written as torch ops
in FLOAT64 — `_tail_torch` (the reference's `rslo/models/odom_pred.py:210-313`, `rslo/layers/confidence.py:23-34`,
`rslo/data/dataset.py:121-208`) and `_loss_tail_eager` (`rslo/models/voxel_odom_net.py:727-795`,
`rslo/core/losses.py:155-197`): every output and every gradient, tolerance 2e-5 of the tensor's max magnitude."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def net(cuda):
    import rslo_b200
    n, _ = rslo_b200.build_network(testing=False, seed=7)
    return n.cuda()


def _rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _tail_inputs(B, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    tq32 = torch.zeros(B, H, W, 32)
    tq32[..., :3] = torch.randn(B, H, W, 3, generator=g) * 0.5
    tq32[..., 3:7] = torch.randn(B, H, W, 4, generator=g) * 0.3 + torch.tensor([1.0, 0, 0, 0])
    tq32[..., 7:] = torch.randn(B, H, W, 25, generator=g)                 # garbage in the unused channels
    tl32 = torch.randn(B, H, W, 32, generator=g) * 2
    rl32 = torch.randn(B, H, W, 32, generator=g) * 2
    py0 = torch.randn(B, H // 4, W // 4, 32, generator=g)
    py1 = torch.randn(B, H // 2, W // 2, 32, generator=g)
    mask = (torch.rand(B, H, W, generator=g) < 0.3).float()
    mask[:, : H // 3] = 0                                                # empty region: max-pool cascade sees zeros
    return [t.cuda() for t in (tq32, tl32, rl32, py0, py1, mask)]


def _weights_like(ts, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(t.shape, generator=g).to(t.device) for t in ts]


@pytest.mark.parametrize("B,H,W", [(2, 96, 176), (3, 24, 44)])
def test_head_tail_kernel_matches_float64(net, B, H, W):
    from rslo_b200.layers.pose_tail import head_geometry, head_tail
    head = net.odom_predictor
    tq32, tl32, rl32, py0, py1, mask = _tail_inputs(B, H, W, 5)
    ins = [t.clone().requires_grad_(True) for t in (tq32, tl32, rl32, py0, py1)]
    geom = head_geometry(H, W, head.point_cloud_range)
    t, q, tq_g, tc, rc, pyr = head_tail(*ins, mask, geom)
    own = [t, q, tq_g, tc, rc, pyr[0][0], pyr[1][0], pyr[2][0]]
    own_masks = [pyr[0][1], pyr[1][1], pyr[2][1]]

    d = [t.detach().double().requires_grad_(True) for t in (tq32, tl32, rl32, py0, py1)]
    nchw = lambda x, c: x[..., :c].permute(0, 3, 1, 2)
    ref = head._tail_torch(nchw(d[0], 7), nchw(d[1], 1), nchw(d[2], 1), [nchw(d[3], 7), nchw(d[4], 7)], mask.double().unsqueeze(1))
    rp = ref["pyramid_motion"]
    refs = [ref["translation_preds"][0], ref["rotation_preds"][0], ref["tq_map_g"], ref["t_conf"], ref["r_conf"],
            rp[0][0], rp[1][0], rp[2][0]]
    names = ["t", "q", "tq_map_g", "t_conf", "r_conf", "py0", "py1", "py2"]
    for n, a, r in zip(names, own, refs):
        assert a.shape == r.shape, n
        assert _rel(a, r) < 2e-5, (n, _rel(a, r))
    for i, (a, r) in enumerate(zip(own_masks, [rp[0][1], rp[1][1], rp[2][1]])):
        assert a.shape == r.shape and _rel(a, r) < 2e-5, ("mask", i, _rel(a, r))
    ws = _weights_like(own, 9)
    ws[0] *= 50; ws[1] *= 50                      # the pose carries the real training signal
    sum((a * w).sum() for a, w in zip(own, ws)).backward()
    sum((r * w.double()).sum() for r, w in zip(refs, ws)).backward()
    for n, a, r in zip(["tq32", "tl32", "rl32", "py0_32", "py1_32"], ins, d):
        assert _rel(a.grad, r.grad) < 2e-5, (n, _rel(a.grad, r.grad))
        if n == "tq32":
            assert float(a.grad[..., 7:].abs().max()) == 0.0


@pytest.mark.parametrize("identity_pose,B", [(False, 2), (True, 1), (False, 3)])
def test_loss_tail_kernel_matches_float64(net, identity_pose, B):
    from rslo_b200.layers.pose_tail import loss_geometry, loss_tail
    H, W = 96, 176
    g = torch.Generator().manual_seed(11 + B)
    T = (torch.randn(B, 3, generator=g) * 0.5).cuda()
    q = torch.nn.functional.normalize(torch.randn(B, 4, generator=g) * 0.2 + torch.tensor([1.0, 0, 0, 0]), dim=1).cuda()
    if B == 3:
        q[2] = torch.nn.functional.normalize(torch.tensor([0.05, 0.9, 0.3, 0.2]), dim=0).cuda()   # trace < 0 branch
    pyr = []
    for s in (4, 2, 1):
        pyr.append([torch.randn(B, 7, H // s, W // s, generator=g).cuda(), torch.rand(B, 2, H // s, W // s, generator=g).cuda() ** 4])
    ang = 0.02 * torch.randn(B, 3, generator=g)
    res_r = torch.matrix_exp(torch.stack([torch.tensor([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]]) for a in ang])).cuda()
    res_t = (0.1 * torch.randn(B, 3, generator=g)).cuda()
    net._translation_loss.alpha.data.fill_(0.3)
    net._rotation_loss.alpha.data.fill_(-2.5)
    mods = (net._translation_loss, net._rotation_loss, net._pyramid_translation_loss, net._pyramid_rotation_loss)
    Tg, qg = T.clone().requires_grad_(True), q.clone().requires_grad_(True)
    pg = [[p.clone().requires_grad_(True), m] for p, m in pyr]
    for m in mods:
        m.alpha.grad = None
    geom = loss_geometry(H, W, net.odom_predictor.point_cloud_range)
    l8, tq_map = loss_tail(Tg, qg, pg, res_r, res_t, [m.alpha for m in mods], [m._loss_weight for m in mods],
                           identity_pose, geom)
    assert l8.shape == (8,)
    own = [l8[i] for i in range(8)]
    coef = [1.0, 0.7, 0.125, 0.25, 0.5, 0.125, 0.25, 0.5]
    (l8 * torch.tensor(coef, device=l8.device)).sum().backward()
    own_alpha = [net._translation_loss.alpha.grad.clone(), net._rotation_loss.alpha.grad.clone()]

    n64 = copy.deepcopy(net).double()
    for m in (n64._translation_loss, n64._rotation_loss):
        m.alpha.grad = None
    Td, qd = T.double().requires_grad_(True), q.double().requires_grad_(True)
    pd = [[p.double().requires_grad_(True), m.double()] for p, m in pyr]
    flat = []
    for p, m in pd:
        flat += [p, m]
    ref = n64._loss_tail_eager(Td, qd, flat, res_r.double(), res_t.double(), identity_pose)
    for i, (o, r) in enumerate(zip(own, ref[:8])):
        assert _rel(o, r) < 2e-5, (i, float(o), float(r))
    assert _rel(tq_map, ref[8]) < 2e-5
    sum(c * o.sum() for c, o in zip(coef, ref[:8])).backward()
    assert _rel(Tg.grad, Td.grad) < 2e-5 and _rel(qg.grad, qd.grad) < 2e-5
    for (a, _), (r, _) in zip(pg, pd):
        assert _rel(a.grad, r.grad) < 2e-5
    assert _rel(own_alpha[0], n64._translation_loss.alpha.grad) < 2e-5
    assert _rel(own_alpha[1], n64._rotation_loss.alpha.grad) < 2e-5
    net._translation_loss.alpha.data.fill_(0.0)
