"""CPU stand-ins with the spconv 1.x module surface, built on the oracle's index tables, so that the
REFERENCE's own `rslo/models/middle.py` and builders run in this container (oracle/ref_shim.py).
Also the CPU-capable stand-in for the reference's chamfer module.  TEST INFRASTRUCTURE ONLY."""
import math

import numpy as np
import torch
from torch import nn

from . import native
from .sparse import gather_conv


class SparseConvTensor:
    def __init__(self, features, indices, spatial_shape, batch_size):
        self.features = features
        self.indices = indices
        self.spatial_shape = [int(s) for s in spatial_shape]
        self.batch_size = batch_size
        self.indice_dict = {}

    def dense(self):
        D, H, W = self.spatial_shape
        C = self.features.shape[1]
        out = self.features.new_zeros((self.batch_size, C, D, H, W))
        idx = self.indices.long()
        out[idx[:, 0], :, idx[:, 1], idx[:, 2], idx[:, 3]] = self.features
        return out


def _triple(v):
    return tuple(int(x) for x in v) if isinstance(v, (list, tuple)) else (int(v),) * 3


class _Conv(nn.Module):
    def __init__(self, cin, cout, kernel_size=3, stride=1, padding=0, bias=True, indice_key=None, subm=False,
                 inverse=False, **kw):
        super().__init__()
        self.ks, self.st, self.pd = _triple(kernel_size), _triple(stride), _triple(padding)
        self.subm, self.inverse, self.indice_key = subm, inverse, indice_key
        self.weight = nn.Parameter(torch.empty(*self.ks, cin, cout))
        self.bias = nn.Parameter(torch.empty(cout)) if bias else None
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if bias:
            fan_in, _ = nn.init._calculate_fan_in_and_fan_out(self.weight)
            nn.init.uniform_(self.bias, -1 / math.sqrt(fan_in), 1 / math.sqrt(fan_in))

    def forward(self, x):
        co = x.indices[:, 1:].cpu().numpy().astype(np.int32)
        out = SparseConvTensor(None, x.indices, x.spatial_shape, x.batch_size)
        out.indice_dict = x.indice_dict
        if self.inverse:
            nbr_inv, in_indices, in_shape = x.indice_dict[self.indice_key][1:]
            table = nbr_inv
            out.indices, out.spatial_shape = in_indices, in_shape
        elif self.subm:
            ent = x.indice_dict.get(self.indice_key)
            if ent is None:
                ent = (native.subm_table(co, x.spatial_shape, self.ks),)
                if self.indice_key:
                    x.indice_dict[self.indice_key] = ent
            table = ent[0]
        else:
            oc, oshape, nbr, nbr_inv = native.strided_table(co, x.spatial_shape, self.ks, self.st, self.pd)
            if self.indice_key:
                x.indice_dict[self.indice_key] = (nbr, nbr_inv, x.indices, x.spatial_shape)
            table = nbr
            out.indices = torch.cat([torch.zeros(len(oc), 1, dtype=torch.int32), torch.from_numpy(oc)], 1)
            out.spatial_shape = oshape
        out.features = gather_conv(x.features, table, self.weight, self.bias)
        return out


class SubMConv3d(_Conv):
    def __init__(self, cin, cout, kernel_size, stride=1, padding=0, bias=True, indice_key=None, **kw):
        super().__init__(cin, cout, kernel_size, stride, padding, bias, indice_key, subm=True)


class SparseConv3d(_Conv):
    def __init__(self, cin, cout, kernel_size, stride=1, padding=0, bias=True, indice_key=None, **kw):
        super().__init__(cin, cout, kernel_size, stride, padding, bias, indice_key)


class SparseInverseConv3d(_Conv):
    def __init__(self, cin, cout, kernel_size, indice_key=None, bias=True, **kw):
        super().__init__(cin, cout, kernel_size, 1, 0, bias, indice_key, inverse=True)


class SparseSequential(nn.Sequential):
    def forward(self, x):
        for m in self._modules.values():
            if isinstance(m, _Conv):
                x = m(x)
            elif isinstance(x, SparseConvTensor):
                if x.indices.shape[0] != 0:
                    x.features = m(x.features)
            else:
                x = m(x)
        return x


class VoxelGenerator:
    """spconv.utils.VoxelGenerator with the kwargs rslo/builder/voxel_builder.py:83-94 passes."""

    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels=20000, full_mean=False,
                 block_filtering=False, block_factor=1, block_size=8, height_threshold=0.2, **kw):
        self.voxel_size = np.array(voxel_size, dtype=np.float32)
        self.point_cloud_range = np.array(point_cloud_range, dtype=np.float32)
        self.max_num_points, self.max_voxels = max_num_points, max_voxels
        self.block_factor, self.block_size, self.height_threshold = block_factor, block_size, height_threshold

    def generate(self, points, max_voxels=None):
        r = native.voxelize(points, self.voxel_size, self.point_cloud_range, self.max_num_points,
                            max_voxels or self.max_voxels, self.block_factor, self.block_size,
                            self.height_threshold)
        return r["voxels"], r["coordinates"], r["num_points_per_voxel"]


class OneDirectionChamferDistanceWithIdx(nn.Module):
    """CPU stand-in for thirdparty/chamfer_distance/chamfer_distance.py:244-246 (the original raises
    on CPU tensors, :174-176).  idx is a constant w.r.t. autograd; dist's gradient is unused on the
    hot path."""

    def forward(self, xyz1, xyz2):
        ds, ids = [], []
        for b in range(xyz1.shape[0]):
            d, i = native.nn(xyz1[b].detach().cpu().numpy(), xyz2[b].detach().cpu().numpy(), fused=True)
            ds.append(torch.from_numpy(d))
            ids.append(torch.from_numpy(i))
        return torch.stack(ds), torch.stack(ids)
