"""Seeded synthetic KITTI-shaped LiDAR scans.

An HDL-64E-shaped sensor (64 beams, elevation +2°..-24.8°, ``n_az`` azimuth steps) is ray-cast
against a ground plane plus random axis-aligned boxes.  The output row layout is the one the
reference dataset hands to the voxeliser (`rslo/data/kitti_dataset_hdf5.py:253-261`):
``[x, y, z, intensity, nx, ny, nz]`` float32, in beam-major scan order, with normals whose
components are exactly (0,0,±1) zeroed as the dataset does (`:261`).

Pure numpy; used by tests, ``bench.py`` and ``__graft_entry__.smoke()``.
"""
import numpy as np

SENSOR_HEIGHT = 1.73


def make_scene(seed=0, n_boxes=40, extent=70.0):
    """Random axis-aligned boxes (walls / cars / poles) standing on the ground plane."""
    rng = np.random.default_rng(seed)
    cx = rng.uniform(-extent, extent, n_boxes)
    cy = rng.uniform(-0.5 * extent, 0.5 * extent, n_boxes)
    # keep the sensor's immediate surroundings free
    near = (np.abs(cx) < 4.0) & (np.abs(cy) < 4.0)
    cx[near] += 8.0
    sx = rng.uniform(0.5, 12.0, n_boxes)
    sy = rng.uniform(0.5, 12.0, n_boxes)
    h = rng.uniform(0.8, 4.5, n_boxes)
    lo = np.stack([cx - sx / 2, cy - sy / 2, np.full(n_boxes, -SENSOR_HEIGHT)], 1)
    hi = np.stack([cx + sx / 2, cy + sy / 2, -SENSOR_HEIGHT + h], 1)
    return lo.astype(np.float64), hi.astype(np.float64)


def _rot_zyx(yaw, pitch, roll):
    cy, sy = np.cos(yaw), np.sin(yaw)
    cp, sp = np.cos(pitch), np.sin(pitch)
    cr, sr = np.cos(roll), np.sin(roll)
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1.0]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    return Rz @ Ry @ Rx


def cast_scan(scene, R=None, t=None, n_beams=64, n_az=1875, noise_seed=0,
              range_sigma=0.02, rmin=2.0, rmax=80.0):
    """Ray-cast one scan from sensor pose (R, t) in the scene frame.

    Returns points [P,7] float32 in the SENSOR frame (misses dropped).
    """
    lo, hi = scene
    R = np.eye(3) if R is None else np.asarray(R, np.float64)
    t = np.zeros(3) if t is None else np.asarray(t, np.float64)
    elev = np.deg2rad(np.linspace(2.0, -24.8, n_beams))
    az = np.linspace(-np.pi, np.pi, n_az, endpoint=False)
    ce, se = np.cos(elev)[:, None], np.sin(elev)[:, None]
    d_s = np.stack([ce * np.cos(az)[None], ce * np.sin(az)[None],
                    np.broadcast_to(se, (n_beams, n_az))], -1).reshape(-1, 3)
    d_w = d_s @ R.T                       # ray directions in the scene frame
    o = t                                 # ray origin
    P = d_w.shape[0]
    best = np.full(P, np.inf)
    normal = np.zeros((P, 3))
    # ground plane z = -SENSOR_HEIGHT
    with np.errstate(divide="ignore", invalid="ignore"):
        tg = (-SENSOR_HEIGHT - o[2]) / d_w[:, 2]
    ok = (d_w[:, 2] < 0) & (tg > 0)
    best[ok] = tg[ok]
    normal[ok] = (0, 0, 1.0)
    # boxes: slab test, vectorised over rays, looped over boxes
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / d_w
    for b in range(lo.shape[0]):
        t0 = (lo[b] - o) * inv
        t1 = (hi[b] - o) * inv
        tn = np.minimum(t0, t1)
        tf = np.maximum(t0, t1)
        tnear = tn.max(1)
        tfar = tf.min(1)
        hit = (tnear < tfar) & (tnear > 0) & (tnear < best)
        if not hit.any():
            continue
        ax = tn[hit].argmax(1)
        n = np.zeros((hit.sum(), 3))
        n[np.arange(len(ax)), ax] = -np.sign(d_w[hit, ax])
        best[hit] = tnear[hit]
        normal[hit] = n
    rng = np.random.default_rng(noise_seed)
    rngs = best + rng.normal(0.0, range_sigma, P)
    inten = rng.uniform(0.0, 1.0, P)
    keep = np.isfinite(best) & (rngs > rmin) & (rngs < rmax)
    pts_s = d_s[keep] * rngs[keep, None]            # sensor frame
    n_s = normal[keep] @ R                          # scene -> sensor frame (R^T n)
    # orient towards the sensor
    flip = (n_s * pts_s).sum(1) > 0
    n_s[flip] *= -1
    out = np.concatenate([pts_s, inten[keep, None], n_s], 1).astype(np.float32)
    nrm = out[:, 4:7]
    nrm[np.abs(nrm) == np.array([0, 0, 1], np.float32)] = 0   # kitti_dataset_hdf5.py:261
    return np.ascontiguousarray(out)


def make_pair(seed=0, n_beams=64, n_az=1875, delta_t=(1.0, 0.02, 0.0),
              delta_ypr_deg=(1.0, 0.1, 0.1)):
    """Two scans of the same scene: frame A at the origin, frame B at pose A∘Δ.

    Returns (points_a [Pa,7], points_b [Pb,7], gt) with gt = (R_ab [3,3], t_ab [3]) such that
    p_a = R_ab p_b + t_ab.
    """
    scene = make_scene(seed)
    Rd = _rot_zyx(*np.deg2rad(np.asarray(delta_ypr_deg, np.float64)))
    td = np.asarray(delta_t, np.float64)
    a = cast_scan(scene, None, None, n_beams, n_az, noise_seed=2 * seed)
    b = cast_scan(scene, Rd, td, n_beams, n_az, noise_seed=2 * seed + 1)
    return a, b, (Rd, td)
