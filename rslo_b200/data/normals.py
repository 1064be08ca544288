"""Normal estimation for raw scans on the device (csrc/normals.cu), mirroring `estimate_normal` of the reference's
offline preprocessing script (`script/create_hdf5.py:130-147`, called at `:322`): open3d's
`estimate_normals(KDTreeSearchParamHybrid(radius=0.6, max_nn=30))` followed by
`orient_normals_towards_camera_location((0,0,0))`.  With it the `[P,7]` input rows (x,y,z,i,nx,ny,nz) of the path can
be produced from a raw KITTI `.bin` scan on the GPU, without the open3d dependency and the per-scan HDF5 detour
(~0.6 s of CPU per 120k-point scan there; one ~2 ms call here)."""
import ctypes as C

import numpy as np
import torch

from .._lib import check, lib, ptr, stream, workspace


def estimate_normal(points, radius=0.6, max_nn=30, camera_location=(0, 0, 0)):
    """points: [N, >=3] float32 tensor (any device) or numpy array -> normals [N,3] float32 on the GPU"""
    t = torch.as_tensor(np.ascontiguousarray(points, dtype=np.float32)) if not torch.is_tensor(points) else points
    t = t.to(device="cuda", dtype=torch.float32)
    assert t.dim() == 2 and t.shape[1] >= 3 and t.stride(1) == 1
    n, ld = t.shape[0], t.stride(0)
    out = torch.empty((n, 3), dtype=torch.float32, device=t.device)
    nb = lib.rslo_estimate_normals_workspace_bytes(n)
    ws = workspace(nb, "normals")
    cam = (C.c_float * 3)(*[float(v) for v in camera_location])
    check(lib.rslo_estimate_normals(ptr(t), ld, n, float(radius), int(max_nn), cam, ptr(out), ptr(ws), ws.numel(), stream()),
          "rslo_estimate_normals")
    return out


def points_with_normals(lidar_points, radius=0.6, max_nn=30):
    """raw KITTI scan [N,4] (x,y,z,intensity) -> the path's [N,7] rows (x,y,z,i,nx,ny,nz) with exact +-(0,0,1) normals
    zeroed as the dataset does (`rslo/data/kitti_dataset_hdf5.py:245-261`)."""
    p = torch.as_tensor(np.ascontiguousarray(lidar_points, dtype=np.float32)) if not torch.is_tensor(lidar_points) else lidar_points
    p = p.to(device="cuda", dtype=torch.float32).contiguous()
    nrm = estimate_normal(p, radius, max_nn)
    up = (nrm[:, 0] == 0) & (nrm[:, 1] == 0) & (nrm[:, 2].abs() == 1)
    nrm = torch.where(up[:, None], torch.zeros_like(nrm), nrm)
    return torch.cat([p[:, :4], nrm], dim=1)
