"""TEST INFRASTRUCTURE ONLY - numpy brute-force restatement of the normal estimation the reference does offline with
open3d (`/root/reference/script/create_hdf5.py:130-147`: `estimate_normals(KDTreeSearchParamHybrid(radius=0.6,
max_nn=30))`, `orient_normals_towards_camera_location`).

PARITY UNPINNED: open3d (freeze.yml pins 0.9.0) is not installed here and the reference holds no normals fixture.
Restated from open3d's published algorithm (geometry/EstimateNormals.cpp): hybrid search = the <= max_nn nearest
neighbours within radius, query point included; >= 3 neighbours -> covariance from cumulants in double ->
eigenvector of the smallest eigenvalue, else (0,0,1); orientation flips normals with n . (camera - p) < 0.
Ties between equidistant neighbours are broken by index here (a KD-tree breaks them by traversal order)."""
import numpy as np


def estimate_normals(points, radius=0.6, max_nn=30, camera=(0.0, 0.0, 0.0)):
    p = np.asarray(points, dtype=np.float32)[:, :3]
    pd = p.astype(np.float64)
    n = len(p)
    out = np.zeros((n, 3), dtype=np.float64)
    lam = np.zeros((n, 3))
    cam = np.asarray(camera, dtype=np.float64)
    r2 = np.float32(radius) * np.float32(radius)
    for i in range(n):
        d = p - p[i]
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]          # float32, the kernel's order
        idx = np.nonzero(d2 <= r2)[0]
        idx = idx[np.lexsort((idx, d2[idx]))][:max_nn]
        nv = np.array([0.0, 0.0, 1.0])
        if len(idx) >= 3:
            q = pd[idx]
            mean = q.mean(0)
            cov = (q[:, :, None] * q[:, None, :]).mean(0) - np.outer(mean, mean)
            w, v = np.linalg.eigh(cov)
            nv = v[:, 0]
            lam[i] = w
        if nv @ (cam - pd[i]) < 0:
            nv = -nv
        out[i] = nv
    return out, lam
