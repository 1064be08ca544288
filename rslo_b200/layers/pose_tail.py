"""The two tails of the dense head as single fused kernels (csrc/pose_tail.cu), forward and backward.

`head_tail`  : `rslo/models/odom_pred.py:226-313` after the convolutions — q normalisation, masked softmax
               confidences (T = 1 and the detached T = 20), local -> global (t,q) (`rslo/data/dataset.py:121-208`),
               confidence vote (`odom_pred.py:347-357`), pyramid masks and masked pyramid predictions.
`loss_tail`  : `rslo/models/voxel_odom_net.py:727-795` — pseudo labels, target (t,q) maps (`dataset.py:52-116`),
               AdaptiveWeightedL2 on the pose and the three pyramid levels (`rslo/core/losses.py:155-197`).
Each replaces ~60 eager torch launches per pass with one launch.
"""
import numpy as np
import torch

from .._lib import TqGeom, check, lib, ptr, stream
from .. import kernels as K


def head_geometry(H, W, pc_range):
    """anchor constants exactly as `rslo_b200.data.dataset.cell_anchors` computes them (float32 tensor arithmetic)"""
    pc = torch.as_tensor(np.asarray(pc_range), dtype=torch.float32)
    grid = torch.tensor([W, H, 1], dtype=torch.float32)
    vs = (pc[3:] - pc[:3]) / grid
    ox = (0 - pc[0]) / (pc[3] - pc[0]) * grid[0]
    oy = (pc[4] - 0) / (pc[4] - pc[1]) * grid[1]
    oz = (0 - pc[2]) / (pc[5] - pc[2]) * grid[2]
    return TqGeom(int(H), int(W), float(ox), float(oy), float(oz), float(vs[0]), float(vs[1]), float(vs[2]))


def loss_geometry(H, W, pc_range):
    """anchor constants as `UnVoxelOdomNetICP3.gen_tq_maps` computes them (numpy, then float32 tensor arithmetic)"""
    pc = np.asarray(pc_range)
    grid = np.array([W, H, 1])
    vs = (pc[3:] - pc[0:3]) / grid
    ox = (0 - pc[0]) / (pc[3] - pc[0]) * grid[0]
    oy = (pc[4] - 0) / (pc[4] - pc[1]) * grid[1]
    oz = (0 - pc[2]) / (pc[5] - pc[2]) * grid[2]
    f = lambda v: float(np.float32(v))
    return TqGeom(int(H), int(W), f(ox), f(oy), f(oz), f(vs[0]), f(vs[1]), f(vs[2]))


class _HeadTailFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tq32, tl32, rl32, py0_32, py1_32, mask, geom):
        B, H, W, _ = tq32.shape
        dev = tq32.device
        f = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        pose_t, pose_q = f(B, 3), f(B, 4)
        tq_g, t_conf, r_conf = f(B, 7, H, W), f(B, 1, H, W), f(B, 1, H, W)
        pm2p, pm2m = f(B, 7, H, W), f(B, 2, H, W)
        pm1p, pm1m = f(B, 7, H // 2, W // 2), f(B, 2, H // 2, W // 2)
        pm0p, pm0m = f(B, 7, H // 4, W // 4), f(B, 2, H // 4, W // 4)
        occ1, occ0, save = f(B, H // 2, W // 2), f(B, H // 4, W // 4), f(B, 16)
        tq32, tl32, rl32 = tq32.contiguous(), tl32.contiguous(), rl32.contiguous()
        py0_32, py1_32, mask = py0_32.contiguous(), py1_32.contiguous(), mask.contiguous()
        check(lib.rslo_head_tail_forward(ptr(tq32), ptr(tl32), ptr(rl32), ptr(mask), ptr(py0_32), ptr(py1_32), B, geom,
                                         ptr(pose_t), ptr(pose_q), ptr(tq_g), ptr(t_conf), ptr(r_conf), ptr(pm2p), ptr(pm2m),
                                         ptr(pm1p), ptr(pm1m), ptr(pm0p), ptr(pm0m), ptr(occ1), ptr(occ0), ptr(save),
                                         stream()), "rslo_head_tail_forward")
        K._count()
        ctx.geom = geom
        ctx.save_for_backward(tq32, mask, t_conf, r_conf, occ1, occ0, save)
        ctx.shapes = (tuple(py0_32.shape), tuple(py1_32.shape))
        ctx.mark_non_differentiable(pm2m, pm1m, pm0m)
        return pose_t, pose_q, tq_g, t_conf, r_conf, pm0p, pm0m, pm1p, pm1m, pm2p, pm2m

    @staticmethod
    def backward(ctx, g_t, g_q, g_tqg, g_tc, g_rc, g_pm0, _m0, g_pm1, _m1, g_pm2, _m2):
        tq32, mask, t_conf, r_conf, occ1, occ0, save = ctx.saved_tensors
        B = tq32.shape[0]
        c = lambda g: None if g is None else g.contiguous()
        g_t, g_q, g_tqg, g_tc, g_rc, g_pm0, g_pm1, g_pm2 = (c(g) for g in (g_t, g_q, g_tqg, g_tc, g_rc, g_pm0, g_pm1, g_pm2))
        d_tq, d_tl, d_rl = torch.empty_like(tq32), torch.empty_like(tq32), torch.empty_like(tq32)
        d_py0 = torch.empty(ctx.shapes[0], dtype=torch.float32, device=tq32.device)
        d_py1 = torch.empty(ctx.shapes[1], dtype=torch.float32, device=tq32.device)
        check(lib.rslo_head_tail_backward(ptr(tq32), ptr(mask), ptr(t_conf), ptr(r_conf), ptr(occ1), ptr(occ0), ptr(save), B,
                                          ctx.geom, ptr(g_t), ptr(g_q), ptr(g_tqg), ptr(g_tc), ptr(g_rc), ptr(g_pm2),
                                          ptr(g_pm1), ptr(g_pm0), ptr(d_tq), ptr(d_tl), ptr(d_rl), ptr(d_py1), ptr(d_py0),
                                          stream()), "rslo_head_tail_backward")
        K._count()
        return d_tq, d_tl, d_rl, d_py0, d_py1, None, None


def head_tail(tq32, tl32, rl32, py0_32, py1_32, mask, geom):
    """-> (pose_t [B,3], pose_q [B,4], tq_map_g [B,7,H,W], t_conf, r_conf [B,1,H,W],
           pyramid_motion = [[pred0, mask0], [pred1, mask1], [pred2, mask2]])"""
    o = _HeadTailFn.apply(tq32, tl32, rl32, py0_32, py1_32, mask, geom)
    return o[0], o[1], o[2], o[3], o[4], [[o[5], o[6]], [o[7], o[8]], [o[9], o[10]]]


class _LossTailFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, T_pred, q_pred, pm0p, pm0m, pm1p, pm1m, pm2p, pm2m, res_r, res_t, a_t, a_r, a_pt, a_pr, weights,
                identity_pose, geom):
        B = T_pred.shape[0]
        dev = T_pred.device
        H, W = pm2p.shape[2:]
        ts = [t.contiguous() for t in (T_pred, q_pred, pm0p, pm0m, pm1p, pm1m, pm2p, pm2m, res_r, res_t)]
        tq_target = torch.empty((B, 7, H, W), dtype=torch.float32, device=dev)
        save = torch.empty((B, 24), dtype=torch.float32, device=dev)
        losses = torch.empty(8, dtype=torch.float32, device=dev)
        counter = torch.zeros(1, dtype=torch.int32, device=dev)
        w = [float(v) for v in weights]
        check(lib.rslo_loss_tail_forward(*[ptr(t) for t in ts], 1 if identity_pose else 0, B, geom, ptr(a_t), ptr(a_r),
                                         ptr(a_pt), ptr(a_pr), w[0], w[1], w[2], w[3], ptr(tq_target), ptr(save),
                                         ptr(losses), ptr(counter), stream()), "rslo_loss_tail_forward")
        K._count()
        ctx.geom, ctx.w, ctx.B = geom, w, B
        ctx.same_t, ctx.same_r = a_pt is a_t or a_pt.data_ptr() == a_t.data_ptr(), a_pr.data_ptr() == a_r.data_ptr()
        ctx.save_for_backward(*ts[:8], a_t, a_r, a_pt, a_pr, save)
        ctx.mark_non_differentiable(tq_target)
        return losses, tq_target

    @staticmethod
    def backward(ctx, g8, _g_target):
        T_pred, q_pred, pm0p, pm0m, pm1p, pm1m, pm2p, pm2m, a_t, a_r, a_pt, a_pr, save = ctx.saved_tensors
        dev = T_pred.device
        g = g8.contiguous()
        dT, dq = torch.empty_like(T_pred), torch.empty_like(q_pred)
        d0, d1, d2 = torch.empty_like(pm0p), torch.empty_like(pm1p), torch.empty_like(pm2p)
        da = torch.empty(4, dtype=torch.float32, device=dev)
        w = ctx.w
        check(lib.rslo_loss_tail_backward(ptr(T_pred), ptr(q_pred), ptr(pm0p), ptr(pm0m), ptr(pm1p), ptr(pm1m), ptr(pm2p),
                                          ptr(pm2m), ctx.B, ctx.geom, ptr(a_t), ptr(a_r), ptr(a_pt), ptr(a_pr), w[0], w[1],
                                          w[2], w[3], ptr(save), ptr(g), ptr(dT), ptr(dq), ptr(d0), ptr(d1), ptr(d2), ptr(da),
                                          stream()), "rslo_loss_tail_backward")
        K._count()
        return (dT, dq, d0, None, d1, None, d2, None, None, None, da[0:1], da[1:2], da[2:3], da[3:4], None, None, None)


def loss_tail(T_pred, q_pred, pyramid, res_r, res_t, alphas, weights, identity_pose, geom):
    """pyramid = [[pred0, mask0], [pred1, mask1], [pred2, mask2]] (coarse to fine); alphas / weights in the order
    (translation, rotation, pyramid translation, pyramid rotation).
    -> (losses8 = [T, R, pyT_0..2, pyR_0..2] as ONE tensor, tq_map_target): a caller that combines the eight terms with
    one weighted sum keeps the backward pass at one kernel (eight separate outputs cost a zeros + stack per step)"""
    (p0, m0), (p1, m1), (p2, m2) = pyramid
    return _LossTailFn.apply(T_pred, q_pred, p0, m0, p1, m1, p2, m2, res_r.detach(), res_t.detach(), *alphas,
                             tuple(weights), bool(identity_pose), geom)
