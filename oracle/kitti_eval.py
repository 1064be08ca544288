"""TEST INFRASTRUCTURE ONLY - CPU (numpy, float64) restatement of the reference's KITTI sequence evaluation, SURVEY §8
row f-N4:
  odom_to_abs_pose                                  /root/reference/rslo/utils/geometric.py:376-406
  qmult / normalize / rotate_vec_by_q               /root/reference/rslo/utils/pose_utils_np.py:127-163,228-243
  kittiOdomEval.trajectoryDistances / calcSequenceErrors / computeSegmentErr / computeSegmentAvgErr
                                                    /root/reference/rslo/utils/kitti_evaluation.py:42-61,95-145,157-198
Pinned by tests/test_cpu_oracle.py::test_kitti_eval_oracle_matches_reference_live against the reference's own functions
imported through oracle/ref_shim.py.  One piece is PARITY UNPINNED: `tq_to_RT` goes through the `numpy-quaternion`
package (`geometric.py:440`, not installed here); its published as_rotation_matrix formula is restated in
`quat_to_matrix` and cross-checked against scipy's Rotation, and the live test hands the reference this function.
"""
import numpy as np

LENGTHS = [100, 200, 300, 400, 500, 600, 700, 800]


def _normalize(x, eps=1e-6):
    return x / (np.linalg.norm(x, axis=1, keepdims=True) + eps)


def qmult(q1, q2):
    q1s, q1v, q2s, q2v = q1[:, :1], q1[:, 1:], q2[:, :1], q2[:, 1:]
    qs = q1s * q2s - np.sum(q1v * q2v, axis=1, keepdims=True)
    qv = q1v * q2s + q2v * q1s + np.cross(q1v, q2v, axis=1)
    return _normalize(np.concatenate((qs, qv), axis=1))


def rotate_vec_by_q(t, q):
    qs, qv = q[:, :1], q[:, 1:]
    b = np.cross(qv, t, axis=1)
    c = 2 * np.cross(qv, b, axis=1)
    return t + 2 * b * qs + c


def odom_to_abs_pose(odoms):
    odoms = np.asarray(odoms, dtype=np.float64)
    t_prev, r_prev = odoms[0][:3].reshape(1, 3), odoms[0][3:].reshape(1, 4)
    out = [np.array([0, 0, 0, 1, 0, 0, 0], dtype=np.float64).reshape(1, 7)]
    for i in range(1, len(odoms)):
        t_cur, r_cur = odoms[i][:3].reshape(1, 3), odoms[i][3:].reshape(1, 4)
        r_new = qmult(r_prev, r_cur)
        t_prev = t_prev + rotate_vec_by_q(t_cur, r_prev)
        r_prev = r_new
        out.append(np.concatenate([t_prev, r_prev], axis=-1))
    return np.concatenate(out, axis=0)


def quat_to_matrix(q):
    """(w,x,y,z), any norm -> 3x3 (numpy-quaternion as_rotation_matrix)"""
    w, x, y, z = (float(v) for v in q)
    s = 2.0 / (w * w + x * x + y * y + z * z)
    return np.array([[1 - s * (y * y + z * z), s * (x * y - z * w), s * (x * z + y * w)],
                     [s * (x * y + z * w), 1 - s * (x * x + z * z), s * (y * z - x * w)],
                     [s * (x * z - y * w), s * (y * z + x * w), 1 - s * (x * x + y * y)]])


def tq_to_RT(tq, expand=True):
    RT = np.eye(4)
    RT[:3, :3] = quat_to_matrix(tq[3:])
    RT[:3, 3] = tq[:3]
    return RT if expand else RT[:3]


def trajectory_distances(poses):
    dist = [0.0]
    for i in range(len(poses) - 1):
        d = poses[i][:3, 3] - poses[i + 1][:3, 3]
        dist.append(dist[i] + np.sqrt(d[0] ** 2 + d[1] ** 2 + d[2] ** 2))
    return dist


def calc_sequence_errors(poses_result, poses_gt, step_size=10):
    gt = [tq_to_RT(np.asarray(p, dtype=np.float64)) for p in poses_gt]
    res = [tq_to_RT(np.asarray(p, dtype=np.float64)) for p in poses_result]
    dist = trajectory_distances(gt)
    err = []
    for first in range(0, len(gt), step_size):
        for len_ in LENGTHS:
            last = -1
            for i in range(first, len(dist)):
                if dist[i] > dist[first] + len_:
                    last = i
                    break
            if last == -1 or last >= len(res) or first >= len(res):
                continue
            d_gt = np.linalg.inv(gt[first]) @ gt[last]
            d_res = np.linalg.inv(res[first]) @ res[last]
            e = np.linalg.inv(d_res) @ d_gt
            r_err = np.arccos(max(min(0.5 * (e[0, 0] + e[1, 1] + e[2, 2] - 1.0), 1.0), -1.0))
            t_err = np.sqrt(e[0, 3] ** 2 + e[1, 3] ** 2 + e[2, 3] ** 2)
            err.append([first, r_err / len_, t_err / len_, len_, len_ / (0.1 * (last - first + 1.0))])
    return err


def segment_errors(seq_errs):
    seg = {l: [] for l in LENGTHS}
    for e in seq_errs:
        seg[e[3]].append([e[2], e[1]])
    return {l: [float(np.mean(np.asarray(v)[:, 0])), float(np.mean(np.asarray(v)[:, 1]))] for l, v in seg.items() if v}


def segment_avg(avg_segment_errs):
    if not avg_segment_errs:
        return 0, 0
    t = sum(v[0] for v in avg_segment_errs.values()) / len(avg_segment_errs)
    r = sum(v[1] for v in avg_segment_errs.values()) / len(avg_segment_errs)
    return t, r
