"""tq-map geometry (`rslo/data/dataset.py:52-208`): dense per-cell local (t,q) maps <-> one global
(t,q).  Cell (i,j,k) of a [Z,Y,X] map has its anchor at
    x = (j - o_x) vs_x,  y = (-i + o_y) vs_y,  z = (k - o_z) vs_z
with the origin cell o derived from the point-cloud range (`dataset.py:145-146,169-171`)."""
import numpy as np
import torch

from ..utils import pose_utils as tch_p

_GRID_CACHE = {}


def cell_anchors(spatial_size_zyx, pc_range, device, dtype):
    """[Y*X*Z, 3] anchor coordinates, flattened (y, x, z)-major as the reference's meshgrid does."""
    size_z, size_y, size_x = (int(s) for s in spatial_size_zyx)
    key = (size_z, size_y, size_x, tuple(float(v) for v in pc_range), str(device), dtype)
    g = _GRID_CACHE.get(key)
    if g is None:
        pc = torch.as_tensor(np.asarray(pc_range), dtype=dtype, device=device)
        grid_size = torch.tensor([size_x, size_y, size_z], dtype=dtype, device=device)
        voxel_size = (pc[3:] - pc[:3]) / grid_size
        ox = (0 - pc[0]) / (pc[3] - pc[0]) * grid_size[0]
        oy = (pc[4] - 0) / (pc[4] - pc[1]) * grid_size[1]
        oz = (0 - pc[2]) / (pc[5] - pc[2]) * grid_size[2]
        iv, jv, kv = torch.meshgrid(torch.arange(size_y, dtype=dtype, device=device),
                                    torch.arange(size_x, dtype=dtype, device=device),
                                    torch.arange(size_z, dtype=dtype, device=device), indexing="ij")
        xv = (jv - ox) * voxel_size[0]
        yv = (-iv + oy) * voxel_size[1]
        zv = (kv - oz) * voxel_size[2]
        g = torch.stack([xv, yv, zv], dim=-1).reshape(-1, 3)
        _GRID_CACHE[key] = g
    return g


def from_pointwise_local_transformation_tch(tq_map, pc_range, inv_trans_factor=-1):
    """local (t,q) map [B,7,Y,X] -> global (t,q) map [B,7,Y,X]: t_g = R(q_l)(t_l - p) + p,
    q_g = normalize(q_l)   (`dataset.py:121-208`)."""
    assert inv_trans_factor <= 0
    B, _, H, W = tq_map.shape
    xyz = cell_anchors((1, H, W), pc_range, tq_map.device, tq_map.dtype)
    tq = tq_map.permute(0, 2, 3, 1).reshape(-1, 7)
    xyzv = xyz.repeat(B, 1)
    t_l, q_l = tq[:, :3], tq[:, 3:]
    t_g = tch_p.rotate_vec_by_q(t=(t_l - xyzv), q=q_l) + xyzv
    q_g = torch.nn.functional.normalize(q_l.view(B, H, W, 4), dim=-1)
    return torch.cat([t_g.view(B, H, W, 3), q_g], dim=-1).permute(0, 3, 1, 2).contiguous()


def generate_pointwise_local_transformation_tch(tq, spatial_size, origin_loc, voxel_size, inv_trans_factor=-1):
    """one global (t,q) [7] -> local (t,q) map [7,Y,X]: t_l = R(q^-1)(t_g - p) + p (`dataset.py:52-116`)."""
    assert inv_trans_factor <= 0
    device, dtype = tq.device, tq.dtype
    t_g, q_g = tq[:3], tq[3:]
    if len(spatial_size) == 2:
        size_x, size_y = (int(s) for s in spatial_size)
        size_z = 1
    else:
        size_x, size_y, size_z = (int(s) for s in spatial_size)
    iv, jv, kv = torch.meshgrid(torch.arange(size_y, device=device), torch.arange(size_x, device=device),
                                torch.arange(size_z, device=device), indexing="ij")
    xv = (jv - origin_loc[0]) * voxel_size[0]
    yv = (-iv + origin_loc[1]) * voxel_size[1]
    zv = (kv - origin_loc[2]) * voxel_size[2]
    xyzv = torch.stack([xv, yv, zv], dim=-1).reshape(-1, 3).to(dtype=dtype)
    t_l = tch_p.rotate_vec_by_q(t=t_g[None] - xyzv, q=tch_p.qinv(q_g[None]).repeat(xyzv.shape[0], 1)) + xyzv
    t_map = t_l.reshape(size_y, size_x, size_z, 3)
    q_map = torch.ones(size_y, size_x, size_z, 4, dtype=dtype, device=device) * q_g
    return torch.cat([t_map, q_map], dim=-1).permute(3, 2, 0, 1).squeeze()
