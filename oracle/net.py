"""CPU restatement of the whole per-frame-pair hot path (a1-a15): voxelise -> VFE -> sparse encoder ->
masked 2-D head -> vote -> loss stack, written functionally over a state_dict with the reference's
key names.  Plain torch CPU (+ oracle/c for the integer parts); it is what travels to the GPU box
as the checker and as bench.py's cpu_baseline / --impl reference arm.

TEST INFRASTRUCTURE ONLY — rslo_b200/ never imports this.

Each function cites the reference code it follows; it is pinned against the reference's own Python
(imported through oracle/ref_shim.py in the build container) by tests/test_cpu_oracle.py and by the
golden fixtures under tests/golden/ (made by tests/golden/make_golden.py from the REFERENCE net).
The head / loss / tq-map rows are therefore pinned to reference code; the voxeliser, rulebook and
sparse-conv rows restate the un-vendored spconv_plus fork => parity unpinned for those.
"""

import numpy as np
import torch
import torch.nn.functional as F

from . import native, quat, sparse

VS = [0.1, 0.1, 0.2]
RANGE = [-70.4, -38.4, -3.0, 70.4, 38.4, 5.0]
SPARSE_SHAPE = [41, 768, 1408]
BN_EPS_HEAD, BN_MOM_HEAD = 1e-3, 0.01          # odom_pred_base.py:140-141


# deterministic weights shared by every implementation (reference net, oracle, CUDA path)
from rslo_b200.utils.weights import deterministic_fill as fill_weights  # noqa: E402,F401


# ------------------------------------------------------------------------------------------------
# pose helpers (rslo/utils/pose_utils.py:130-142; kornia conversions in oracle/quat.py)
# ------------------------------------------------------------------------------------------------
def rotate_vec_by_q(t, q):
    qs, qv = q[:, :1], q[:, 1:]
    b = torch.cross(qv, t, dim=1)
    c = 2 * torch.cross(qv, b, dim=1)
    b = 2 * b.mul(qs.expand_as(b))
    return t + b + c


def qinv(q):
    return torch.cat([q[:, :1], -q[:, 1:]], dim=1)


def roll(x, shift, dim=-1):
    return torch.roll(x, shifts=shift, dims=dim)


def _cell_centres(size_z, size_y, size_x, pc_range, dtype):
    """dataset.py:139-173: x=(j-ox)*vx, y=(-i+oy)*vy, z=(k-oz)*vz on a (y,x,z) meshgrid."""
    pc = torch.as_tensor(np.asarray(pc_range), dtype=dtype)
    grid = torch.tensor([size_x, size_y, size_z], dtype=dtype)
    vs = (pc[3:] - pc[:3]) / grid
    ox = (0 - pc[0]) / (pc[3] - pc[0]) * grid[0]
    oy = (pc[4] - 0) / (pc[4] - pc[1]) * grid[1]
    oz = (0 - pc[2]) / (pc[5] - pc[2]) * grid[2]
    iv, jv, kv = torch.meshgrid(torch.arange(size_y, dtype=dtype), torch.arange(size_x, dtype=dtype),
                                torch.arange(size_z, dtype=dtype), indexing="ij")
    xv = (jv - ox) * vs[0]
    yv = (-iv + oy) * vs[1]
    zv = (kv - oz) * vs[2]
    return torch.stack([xv, yv, zv], dim=-1).reshape(-1, 3)


def from_pointwise_local_transformation(tq_map, pc_range):
    """dataset.py:121-208."""
    B, _, H, W = tq_map.shape
    flat = tq_map.permute(0, 2, 3, 1).contiguous().view(-1, 7)
    t_l, q_l = flat[:, :3], flat[:, 3:]
    xyzv = torch.cat([_cell_centres(1, H, W, pc_range, tq_map.dtype)] * B, dim=0)
    t_g = rotate_vec_by_q(t_l - xyzv, q_l) + xyzv
    t_map_g = t_g.view(B, H, W, 3)
    q_map_g = F.normalize(q_l.view(B, H, W, 4), dim=-1)
    return torch.cat([t_map_g, q_map_g], dim=-1).permute(0, 3, 1, 2).contiguous()


def generate_pointwise_local_transformation(tq, H, W, pc_range):
    """dataset.py:52-116 as called by gen_tq_maps (voxel_odom_net.py:293-322): global (t,q) ->
    per-cell local map [7,H,W]."""
    t_g, q_g = tq[:3], tq[3:]
    # gen_tq_maps computes voxel_size / origin in float64 numpy from the float32 pc_range
    pc = np.asarray(pc_range)
    grid = np.array([W, H, 1])
    vs = (pc[3:] - pc[0:3]) / grid
    origin = ((0 - pc[0]) / (pc[3] - pc[0]) * grid[0], (pc[4] - 0) / (pc[4] - pc[1]) * grid[1],
              (0 - pc[2]) / (pc[5] - pc[2]) * grid[2])
    iv, jv, kv = torch.meshgrid(torch.arange(H), torch.arange(W), torch.arange(1), indexing="ij")
    xv = (jv - origin[0]) * vs[0]
    yv = (-iv + origin[1]) * vs[1]
    zv = (kv - origin[2]) * vs[2]
    xyzv = torch.stack([xv, yv, zv], dim=-1).reshape(-1, 3).to(dtype=tq.dtype)
    t_l = rotate_vec_by_q(t_g[None] - xyzv, qinv(q_g[None]).repeat(xyzv.shape[0], 1)) + xyzv
    t_map = t_l.reshape(H, W, 1, 3)
    q_map = torch.ones(H, W, 1, 4, dtype=tq.dtype) * q_g
    return torch.cat([t_map, q_map], dim=-1).permute(3, 2, 0, 1).squeeze()


# ------------------------------------------------------------------------------------------------
# a8/a9: head (odom_pred.py:152-361, odom_pred_base.py:155-276, custom_resnet_spc.py:224-298)
# ------------------------------------------------------------------------------------------------
class _BN:
    def __init__(self, sd, training, stats_out=None):
        self.sd, self.training, self.stats_out = sd, training, stats_out

    def __call__(self, x, p):
        sd = self.sd
        return F.batch_norm(x, sd[p + ".running_mean"].clone(), sd[p + ".running_var"].clone(), sd[p + ".weight"],
                            sd[p + ".bias"], self.training, BN_MOM_HEAD, BN_EPS_HEAD)


def _conv(sd, x, p, stride=1, padding=1):
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=stride, padding=padding)


def _seq3(sd, bn, x, p):
    """conv3x3-BN-ReLU, conv3x3-BN-ReLU, conv1x1: tq_map_conv / confidence / pyramid stacks."""
    x = F.relu(bn(_conv(sd, x, p + ".0"), p + ".1"))
    x = F.relu(bn(_conv(sd, x, p + ".3"), p + ".4"))
    return _conv(sd, x, p + ".6", padding=0)


def _softmax_conf(logit, mask, temperature=1.0):
    """confidence.py:23-34."""
    conf = torch.where(mask > 0, logit, torch.full_like(logit, -1000))
    shape = conf.shape
    return F.softmax(conf.reshape(shape[0], shape[1], -1) / temperature, dim=-1).reshape(shape)


def head_forward(sd, bevs, training=False, pc_range=RANGE, prefix="odom_predictor.",
                 layer_nums=(3, 5, 5)):
    """bevs: list of T dense maps [1,128,H,W] -> dict like the reference head's ret_dict."""
    sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    bn = _BN(sd, training)
    x1, x2 = [], []                                          # odom_pred_base.py:305-324
    for i in range(len(bevs)):
        for j in range(i + 1, len(bevs)):
            x1.append(bevs[i])
            x2.append(bevs[j])
    xs = [torch.cat(x1, 0), torch.cat(x2, 0)]
    input_mask = (torch.sum(xs[0], dim=1, keepdim=True) != 0).to(xs[0].dtype)
    x = torch.cat(xs, dim=1)
    ups = []
    for s, nblk in enumerate(layer_nums):
        for b in range(nblk):
            p = f"blocks.{s}.{b}"
            stride = 2 if b == 0 else 1
            out = F.relu(bn(F.conv2d(x, sd[p + ".conv1.conv1.weight"], None, stride, 1), p + ".bn1"))
            out = bn(F.conv2d(out, sd[p + ".conv2.conv1.weight"], None, 1, 1), p + ".bn2")
            res = x
            if b == 0:
                res = bn(F.conv2d(x, sd[p + ".downsample.0.conv1.weight"], None, stride, 0), p + ".downsample.1")
            x = F.relu(out + res)
        ups.append(F.relu(bn(_conv(sd, x, f"skip_blocks.{s}.0"), f"skip_blocks.{s}.1")))
    py_masks, p_mask = [], input_mask
    for _ in range(2):                                       # mask_gen_pools, odom_pred.py:210-216
        p_mask = F.max_pool2d(p_mask, 3, 2, 1)
        py_masks.append(p_mask)
    py_masks.reverse()
    py_preds = []
    for i in range(3):
        x = torch.cat([x, ups[-(i + 1)]], dim=1)
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        x = F.relu(bn(_conv(sd, x, f"deblocks.{i}.1"), f"deblocks.{i}.2"))
        if i < 2:
            py = _seq3(sd, bn, x, f"pyramid_motion_blocks.{i}")
            py_preds.append([py * (py_masks[i] > 0).to(py.dtype), py_masks[i]])
    x_tail = x
    tq_map = _seq3(sd, bn, x, "tq_map_conv")
    q_map = tq_map[:, 3:] / torch.norm(tq_map[:, 3:], dim=1, keepdim=True)
    tq_map = torch.cat([tq_map[:, :3], q_map], dim=1)
    t_logit = _seq3(sd, bn, x_tail, "t_map_conf.conf_model")
    r_logit = _seq3(sd, bn, x_tail, "q_map_conf.conf_model")
    t_conf = _softmax_conf(t_logit, input_mask)
    r_conf = _softmax_conf(r_logit, input_mask)
    tq_map_g = from_pointwise_local_transformation(tq_map, pc_range)
    t = torch.sum(tq_map_g[:, :3] * t_conf, dim=(2, 3)) / (torch.sum(t_conf, dim=(2, 3)) + 1e-12)
    q = torch.sum(tq_map_g[:, 3:] * r_conf, dim=(2, 3)) / (torch.sum(r_conf, dim=(2, 3)) + 1e-12)
    # the reference re-runs the conf stacks on x_tail.detach() at temperature 20 (odom_pred.py:255-258);
    # in training mode that second pass sees the same batch statistics, so the logits are identical
    temp = torch.cat([_softmax_conf(t_logit.detach(), input_mask, 20),
                      _softmax_conf(r_logit.detach(), input_mask, 20)], dim=1).detach()
    pyramid = py_preds + [[tq_map * input_mask, input_mask * temp]]
    for p in range(2, len(pyramid) + 1):
        pyramid[-p][1] = pyramid[-p][1] * F.avg_pool2d(pyramid[-(p - 1)][1], 3, 2, 1)
    q = q / (torch.norm(q, dim=1, keepdim=True) + 1e-12)
    return {"translation_preds": t, "rotation_preds": q, "tq_map_g": tq_map_g * input_mask,
            "pyramid_motion": pyramid, "t_conf": t_conf, "r_conf": r_conf, "input_mask": input_mask}


# ------------------------------------------------------------------------------------------------
# a11/a12: consistency loss (losses.py:301-507) and SVDHead (svd.py:13-64)
# ------------------------------------------------------------------------------------------------
def svd_head(src, tgt, weight):
    """src/tgt [1,3,n], weight [1,n] -> (R^T [1,3,3], -R^T t [1,3])."""
    src_c = src - src.mean(dim=2, keepdim=True)
    tgt_c = tgt - tgt.mean(dim=2, keepdim=True)
    H = torch.matmul(src_c * weight[:, None, :], tgt_c.transpose(2, 1).contiguous())
    u, s, v = torch.svd(H[0])
    r = v @ u.t()
    if torch.det(r) < 0:
        reflect = torch.eye(3, dtype=src.dtype)
        reflect[2, 2] = -1
        v = v @ reflect
        r = v @ u.t()
    R = r[None]
    t = torch.matmul(-R, src.mean(dim=2, keepdim=True)) + tgt.mean(dim=2, keepdim=True)
    Rt = R.transpose(-1, -2)
    return Rt, -(Rt @ t).squeeze(-1)


def span_cov2(p):
    """losses.py:348-363."""
    c = p.clone()
    c[:, 1:2] = c[:, 0:1] + p[:, 1:2]
    c[:, 2:3] = c[:, 1:2] + p[:, 2:3]
    c[:, 3:] = c[:, 3:].clone() / (torch.norm(p[:, 3:], dim=-1, keepdim=True) + 1e-9)
    eigval = torch.zeros(c.shape[0], 9, dtype=c.dtype)
    eigval[:, ::4] = c[:, :3]
    eigval = eigval.reshape(-1, 3, 3)
    eigvec = quat.quaternion_to_rotation_matrix(c[:, 3:])
    return eigvec @ eigval @ eigvec.transpose(-1, -2)


def points_roi(dist, ratio):
    """losses.py:326-334 (dist [1,N])."""
    flat = dist.reshape(-1)
    m, _ = torch.kthvalue(flat, 1 + int(len(flat) * ratio), dim=-1)
    m = torch.max(m, torch.ones_like(m))
    return dist < m


def _cd(a, b):
    d, i = native.nn(a.detach().numpy(), b.detach().numpy(), fused=True)
    return torch.from_numpy(d), torch.from_numpy(i).long()


def consistency_loss(xyz_pred, xyz_target, cov_pred, cov_target, R_pred, normal_pred, icp_iter,
                     penalize_ratio=np.float32(0.97), reg_weight=np.float32(0.005), alpha=0.0):
    """losses.py:337-507; tensors are batched over pairs: [B,N,3], [B,N,7], [B,3,3]."""
    penalize_ratio, reg_weight = float(penalize_ratio), float(reg_weight)
    loss, res_R, res_T = 0, [], []
    for b in range(xyz_pred.shape[0]):
        cp = span_cov2(cov_pred[b])
        ct = span_cov2(cov_target[b])
        diff, idx1 = _cd(xyz_pred[b], xyz_target[b])
        xyz_assoc = xyz_target[b][idx1]
        ct_assoc = ct[idx1]
        diff_vec = xyz_pred[b] - xyz_assoc
        weight = F.cosine_similarity(normal_pred[b], xyz_assoc - xyz_pred[b], dim=-1)[..., None].abs()
        count_mask = points_roi(diff[None], penalize_ratio)
        sel = count_mask.squeeze(0)
        Rb = R_pred[b].detach()
        sigma = cp[sel] + Rb @ ct_assoc[sel] @ Rb.transpose(-1, -2)
        sigma_inv = torch.inverse(sigma)
        dv = diff_vec[sel]
        square_diff = (dv.unsqueeze(-2) @ sigma_inv @ dv[..., None]).squeeze(-1)
        loss = loss + torch.mean(square_diff) + reg_weight * torch.mean(0.5 * torch.log(torch.det(sigma)))
        src = xyz_pred[b][sel].reshape(1, -1, 3).permute(0, 2, 1).detach()
        tgt = xyz_assoc[sel].reshape(1, -1, 3).permute(0, 2, 1).detach()
        wgt = weight.squeeze(-1)[sel].reshape(1, -1).detach()
        res_r_ = torch.eye(3)
        res_t_ = torch.zeros(3)
        for it in range(icp_iter):
            R, t = svd_head(src, tgt, wgt ** 2)
            res_r_ = R @ res_r_
            res_t_ = (R @ res_t_[..., None] + t[..., None]).squeeze(-1)
            if it < icp_iter - 1:
                moved = (torch.matmul(res_r_[:, None], xyz_target[b][None, ..., None])
                         + res_t_[:, None, :, None]).squeeze(0).squeeze(-1).detach()
                od, oi = _cd(xyz_pred[b], moved)
                assoc = moved[oi]
                w2 = F.cosine_similarity(normal_pred[b], assoc - xyz_pred[b], dim=-1)[..., None].abs()
                roi = points_roi(od[None], penalize_ratio).squeeze(0)
                src = xyz_pred[b][roi].reshape(1, -1, 3).permute(0, 2, 1).detach()
                tgt = assoc[roi].reshape(1, -1, 3).permute(0, 2, 1).detach()
                wgt = w2.squeeze(-1)[roi].reshape(1, -1).detach()
        res_R.append(res_r_)
        res_T.append(res_t_)
    res_R = torch.cat(res_R, dim=0)
    res_T = torch.cat(res_T, dim=0)
    loss = loss / xyz_pred.shape[0]
    a = torch.tensor([alpha])
    return (torch.exp(-a) * loss).sum() + a, res_R, res_T       # focal_gamma 0 => focal weight 1


def adaptive_l2(pred, target, alpha, mask=None):
    """losses.py:155-197 with focal_gamma 0."""
    mask = torch.ones_like(target) if mask is None else mask.expand_as(target)
    sq = (pred - target) ** 2 * mask
    dims = list(range(1, pred.dim()))
    loss = torch.sum(sq, dim=dims) / (torch.sum(mask, dim=dims) + 1e-12)
    fw = torch.ones_like(loss)
    fw = fw / (torch.sum(fw) + 1e-12)
    return (fw * (torch.exp(-alpha) * loss)).sum() + alpha


def loss_forward(sd, head, voxel_features, cov_preds, step, icp_iter_cfg=2, pc_range=RANGE,
                 pyloss_exp_w_base=0.5):
    """voxel_odom_net.py:324-376 + 586-798 for the shipped config (warm_flag False, weights 1)."""
    T_pred, q_pred = head["translation_preds"], head["rotation_preds"]
    pts = [p[:, [0, 1, 2, 4, 5, 6]] for p in voxel_features]
    min_len = min(p.shape[0] for p in pts)
    pts = [p[:min_len] for p in pts]
    covs = [c[:min_len] for c in cov_preds]
    p0, p1, c0, c1 = [], [], [], []
    for i in range(len(pts)):
        for j in range(i + 1, len(pts)):
            p0.append(pts[i]); p1.append(pts[j]); c0.append(covs[i]); c1.append(covs[j])
    p0, p1, c0, c1 = torch.stack(p0), torch.stack(p1), torch.stack(c0), torch.stack(c1)
    R_pred = quat.quaternion_to_rotation_matrix(roll(q_pred, -1))
    T_used = T_pred
    if step <= 1500:
        R_pred = torch.stack([torch.eye(3)] * R_pred.shape[0], dim=0)
        T_used = torch.zeros_like(T_pred)
    tgt = (torch.matmul(R_pred[:, None], p1[:, :, :3][..., None]) + T_used[:, None, :, None]).squeeze(-1)
    icp_iter = icp_iter_cfg if step > 1500 else 5
    l, res_r, res_t = consistency_loss(p0[:, :, :3], tgt, c0, c1, R_pred, p0[:, :, 3:].detach(), icp_iter,
                                       alpha=float(sd.get("_consistency_loss.alpha", torch.zeros(1))[0]))
    C_loss = 1.0 * l
    rot_t = quat.rotation_matrix_to_quaternion(res_r @ R_pred.detach())
    rot_t = roll(rot_t, 1)
    rot_t = rot_t * torch.sign(rot_t[:, 0:1])
    trans_t = (res_r @ T_used[..., None].detach() + res_t[..., None]).squeeze(-1)
    pyramid = head["pyramid_motion"]
    H, W = pyramid[-1][0].shape[2:]
    tq = torch.cat([trans_t, rot_t], dim=-1).reshape(-1, 7)
    tq_maps = torch.stack([generate_pointwise_local_transformation(t, H, W, np.asarray(pc_range, np.float32))
                           for t in tq], dim=0)
    a_t, a_r = sd["_translation_loss.alpha"], sd["_rotation_loss.alpha"]
    T_loss = adaptive_l2(T_pred, trans_t, a_t)
    R_loss = adaptive_l2(q_pred, rot_t, a_r)
    py_loss = torch.zeros(1)
    n = len(pyramid)
    for i, (pred, mask) in enumerate(pyramid):
        Tt, Rt = tq_maps[:, :3], tq_maps[:, 3:]
        if Tt.shape != pred[:, :3].shape:
            Tt = F.interpolate(Tt, size=pred.shape[2:], mode="nearest")
            Rt = F.interpolate(Rt, size=pred.shape[2:], mode="nearest")
        lt = adaptive_l2(pred[:, :3], Tt, a_t, mask=mask[:, :1])
        lr = adaptive_l2(pred[:, 3:], Rt, a_r, mask=mask[:, -1:])
        py_loss = py_loss + pyloss_exp_w_base ** (n - i) * (lt + lr)
    loss = T_loss + R_loss + py_loss + C_loss
    return {"loss": loss, "translation_loss": T_loss, "rotation_loss": R_loss, "pyramid_loss": py_loss,
            "C_loss": C_loss, "res_r": res_r, "res_t": res_t, "rotation_targets": rot_t,
            "translation_targets": trans_t}


# ------------------------------------------------------------------------------------------------
# whole path
# ------------------------------------------------------------------------------------------------
def sparse_shape_of(voxel_size):
    """grid (x,y,z) = round(range / voxel_size); sparse shape (z+1, y, x)  (`voxel_builder.py`, `middle.py:111`)"""
    r = np.asarray(RANGE, np.float64)
    g = np.round((r[3:] - r[:3]) / np.asarray(voxel_size, np.float64)).astype(np.int64)
    return [int(g[2]) + 1, int(g[1]), int(g[0])]


def encode_frame(sd, points, training=False, max_voxels=40000, voxel_size=None):
    vs = VS if voxel_size is None else list(voxel_size)
    shape = SPARSE_SHAPE if voxel_size is None else sparse_shape_of(vs)
    vox = native.voxelize(points, vs, RANGE, 10, max_voxels, 1, 8, -1.0)
    feat = sparse.vfe_mean(vox["voxels"], vox["num_points_per_voxel"])
    n = feat.shape[0]
    coors = np.concatenate([np.zeros((n, 1), np.int32), vox["coordinates"]], 1)
    bev, cov, tables = sparse.middle_forward(sd, feat, coors, shape, training=training)
    return feat, bev, cov, vox, tables


def pair_forward(sd, frames, training=False, step=2000, with_loss=None, grads_for=(), voxel_size=None, max_voxels=40000):
    """frames: list of T point arrays [P,7].  Returns dict(pose [B,7], loss terms, maps...).
    `grads_for`: state_dict keys whose d(loss)/d(param) to return (training only).
    `voxel_size` / `max_voxels`: other than the shipped 0.1 x 0.1 x 0.2 m / 40000 (dense-scan stress configuration)."""
    sd = dict(sd)
    for k in grads_for:
        sd[k] = sd[k].clone().requires_grad_(True)
    with_loss = training if with_loss is None else with_loss
    ctx = torch.enable_grad() if grads_for else torch.no_grad()
    with ctx:
        feats, bevs, covs, n_vox = [], [], [], []
        for pts in frames:
            f, bev, cov, vox, _ = encode_frame(sd, pts, training, max_voxels=max_voxels, voxel_size=voxel_size)
            feats.append(f); bevs.append(bev); covs.append(cov); n_vox.append(f.shape[0])
        head = head_forward(sd, bevs, training)
        out = {"pose": torch.cat([head["translation_preds"], head["rotation_preds"]], -1).detach().numpy(),
               "n_voxels": n_vox, "head": head, "voxel_features": feats, "cov": covs, "bev": bevs}
        if with_loss:
            L = loss_forward(sd, head, feats, covs, step)
            out.update({k: (v.detach().numpy() if torch.is_tensor(v) else v) for k, v in L.items()})
            if grads_for:
                gs = torch.autograd.grad(L["loss"].sum(), [sd[k] for k in grads_for], allow_unused=True)
                out["grads"] = {k: (None if g is None else g.numpy()) for k, g in zip(grads_for, gs)}
    return out
