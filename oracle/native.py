"""ctypes bindings of oracle/c/oracle.c (numpy in, numpy out).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build_oracle())
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


_GRID_CACHE = {}


def voxelize(points, voxel_size, pc_range, max_points=10, max_voxels=40000, block_factor=1,
             block_size=8, height_threshold=-1.0):
    """a1 restatement -> dict(voxels [N,max_points,F] f32, coordinates [N,3] i32 (z,y,x),
    num_points_per_voxel [N] i32).  Mirrors `_VoxelGenerator.generate` (voxel_builder.py:48-54)."""
    points = np.ascontiguousarray(points, np.float32)
    P, F = points.shape
    vs = np.asarray(voxel_size, np.float32)
    rg = np.asarray(pc_range, np.float32)
    g = np.round((np.asarray(pc_range, np.float64)[3:] - np.asarray(pc_range, np.float64)[:3])
                 / np.asarray(voxel_size, np.float64)).astype(np.int64)
    gx, gy, gz = int(g[0]), int(g[1]), int(g[2])
    key = (gx, gy, gz)
    grid = _GRID_CACHE.get(key)
    if grid is None:
        grid = np.full(gx * gy * gz, -1, np.int32)
        _GRID_CACHE[key] = grid
    voxels = np.zeros((max_voxels, max_points, F), np.float32)
    coors = np.zeros((max_voxels, 3), np.int32)
    num = np.zeros((max_voxels,), np.int32)
    f = lib().oracle_voxelize
    f.restype = C.c_int
    n = f(_p(points, C.c_float), C.c_int(P), C.c_int(F), _p(vs, C.c_float), _p(rg, C.c_float),
          C.c_int(gx), C.c_int(gy), C.c_int(gz), C.c_int(max_points), C.c_int(max_voxels),
          C.c_int(block_factor), C.c_int(block_size), C.c_float(height_threshold),
          _p(grid, C.c_int32), _p(voxels, C.c_float), _p(coors, C.c_int32), _p(num, C.c_int32))
    return {"voxels": voxels[:n].copy(), "coordinates": coors[:n].copy(),
            "num_points_per_voxel": num[:n].copy()}


def nn(query, target, fused=True):
    """a10 restatement: (dist [n] f32, idx [n] i32) of each query's nearest target."""
    q = np.ascontiguousarray(query, np.float32).reshape(-1, 3)
    t = np.ascontiguousarray(target, np.float32).reshape(-1, 3)
    dist = np.zeros(q.shape[0], np.float32)
    idx = np.zeros(q.shape[0], np.int32)
    f = lib().oracle_nn
    f.restype = None
    f(_p(q, C.c_float), C.c_int(q.shape[0]), _p(t, C.c_float), C.c_int(t.shape[0]),
      C.c_int(1 if fused else 0), _p(dist, C.c_float), _p(idx, C.c_int32))
    return dist, idx


def subm_table(coors, shape, ksize=(3, 3, 3)):
    """a5 (SubMConv3d): nbr [N, K] int32."""
    coors = np.ascontiguousarray(coors, np.int32)
    n = coors.shape[0]
    K = int(np.prod(ksize))
    nbr = np.empty((n, K), np.int32)
    shp = np.asarray(shape, np.int32)
    ks = np.asarray(ksize, np.int32)
    f = lib().oracle_subm_table
    f.restype = None
    f(_p(coors, C.c_int32), C.c_int(n), _p(shp, C.c_int32), _p(ks, C.c_int32), _p(nbr, C.c_int32))
    return nbr


def strided_table(coors, shape, ksize, stride, pad):
    """a5 (SparseConv3d): (out_coors [M,3], out_shape, nbr [M,K], nbr_inv [N,K])."""
    coors = np.ascontiguousarray(coors, np.int32)
    n = coors.shape[0]
    K = int(np.prod(ksize))
    shp = np.asarray(shape, np.int32)
    ks = np.asarray(ksize, np.int32)
    st = np.asarray(stride, np.int32)
    pd = np.asarray(pad, np.int32)
    out_shape = np.zeros(3, np.int32)
    cap = max(n * K, 1)
    out_coors = np.empty((cap, 3), np.int32)
    nbr = np.empty((cap, K), np.int32)
    nbr_inv = np.empty((max(n, 1), K), np.int32)
    f = lib().oracle_strided_table
    f.restype = C.c_int
    m = f(_p(coors, C.c_int32), C.c_int(n), _p(shp, C.c_int32), _p(ks, C.c_int32), _p(st, C.c_int32),
          _p(pd, C.c_int32), _p(out_shape, C.c_int32), _p(out_coors, C.c_int32), _p(nbr, C.c_int32),
          _p(nbr_inv, C.c_int32))
    return out_coors[:m].copy(), out_shape.tolist(), nbr[:m].copy(), nbr_inv[:n].copy()
