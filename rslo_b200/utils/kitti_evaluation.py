"""KITTI odometry sequence evaluation on the device (csrc/kitti_eval.cu), mirroring the reference's
`rslo/utils/kitti_evaluation.py:kittiOdomEval` and `rslo/utils/geometric.py:odom_to_abs_pose` (same names, argument
meaning and return structures), as called from `rslo/data/kitti_dataset_hdf5.py:346-365`.

The reference runs these as pure-Python loops on rank 0 while the other ranks wait at a barrier
(`train_hdf5.py:872-886`); here the pose chain, the trajectory distances and the 8 segment errors of every 10th start
frame are three kernel launches in float64.  Inputs may be numpy arrays or tensors on any device; computation happens
on the current CUDA device.
"""
import numpy as np
import torch

from .._lib import check, lib, ptr, stream

LENGTHS = [100, 200, 300, 400, 500, 600, 700, 800]


def _dev64(x):
    t = torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x)
    return t.to(device="cuda", dtype=torch.float64).reshape(-1, 7).contiguous()


def odom_to_abs_pose_device(odoms, odoms_gt=None):
    """-> (abs [N,7], abs_gt [N,7] or None, dist_gt [N] or None) float64 device tensors"""
    a = _dev64(odoms)
    n = a.shape[0]
    b = _dev64(odoms_gt) if odoms_gt is not None else None
    assert b is None or b.shape[0] == n, "one launch chains two sequences of the same length"
    abs_a = torch.empty_like(a)
    abs_b = torch.empty_like(b) if b is not None else None
    dist = torch.empty(n, dtype=torch.float64, device=a.device) if b is not None else None
    check(lib.rslo_odom_to_abs_pose(ptr(a), ptr(b), n, ptr(abs_a), ptr(abs_b), ptr(dist), stream()), "rslo_odom_to_abs_pose")
    return abs_a, abs_b, dist


def odom_to_abs_pose(odoms):
    """`geometric.odom_to_abs_pose`: [N,7] relative (t, q wxyz) -> [N,7] absolute poses (numpy, like the reference)"""
    return odom_to_abs_pose_device(odoms)[0].cpu().numpy()


class kittiOdomEval:
    def __init__(self):
        self.lengths = list(LENGTHS)
        self.num_lengths = len(self.lengths)
        self.step_size = 10
        self.max_speed = 0
        self.distance = 0.0

    def _errors_device(self, poses_result, poses_gt):
        res, gt = _dev64(poses_result), _dev64(poses_gt)
        n_gt = gt.shape[0]
        # cumulative distances of the ground truth: the chain kernel computes them from RELATIVE poses; for absolute
        # poses given directly they are a cumulative sum of consecutive position differences, in the same order
        d = (gt[:-1, :3] - gt[1:, :3])
        step = torch.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2])
        dist = torch.cat([torch.zeros(1, dtype=torch.float64, device=gt.device), torch.cumsum(step, 0)])
        rows = ((n_gt + self.step_size - 1) // self.step_size) * 8
        err = torch.empty((rows, 5), dtype=torch.float64, device=gt.device)
        valid = torch.empty(rows, dtype=torch.int32, device=gt.device)
        check(lib.rslo_kitti_sequence_errors(ptr(res), res.shape[0], ptr(gt), n_gt, ptr(dist), self.step_size, ptr(err),
                                             ptr(valid), stream()), "rslo_kitti_sequence_errors")
        return err, valid, dist

    def calcSequenceErrors(self, poses_result, poses_gt):
        """-> list of [first_frame, r_err/len, t_err/len, len, speed] (`kitti_evaluation.py:95-145`)"""
        err, valid, dist = self._errors_device(poses_result, poses_gt)
        e = err[valid.bool()].cpu().numpy()
        self.distance = float(dist[-1])
        self.max_speed = float(e[:, 4].max()) if len(e) else 0
        return [[int(r[0]), float(r[1]), float(r[2]), int(r[3]), float(r[4])] for r in e]

    def computeOverallErr(self, seq_err):
        n = len(seq_err)
        return sum(e[2] for e in seq_err) / n, sum(e[1] for e in seq_err) / n

    def computeSegmentErr(self, seq_errs, return_seg_err=False):
        seg = {l: [] for l in self.lengths}
        for e in seq_errs:
            seg[e[3]].append([e[2], e[1]])
        avg = {l: [float(np.mean(np.asarray(v)[:, 0])), float(np.mean(np.asarray(v)[:, 1]))] for l, v in seg.items() if v}
        return (avg, seg) if return_seg_err else avg

    def computeSegmentAvgErr(self, segment_errs):
        if len(segment_errs) == 0:
            return 0, 0
        t = sum(v[0] for v in segment_errs.values())
        r = sum(v[1] for v in segment_errs.values())
        return t / len(segment_errs), r / len(segment_errs)

    def evaluate_odometry(self, odoms_pred, odoms_gt):
        """The whole of `kitti_dataset_hdf5.py:346-365` in one go: relative predictions / ground truth [N,7] ->
        {"kitti_error": per-length [t, r], "kitti_avg_error": (t, r)}; the chain, the distances and the segment errors
        never leave the device until the (<= 3640-row) error table is read."""
        abs_p, abs_g, _ = odom_to_abs_pose_device(odoms_pred, odoms_gt)
        errs = self.calcSequenceErrors(abs_p, abs_g)
        seg = self.computeSegmentErr(errs)
        return {"kitti_error": seg, "kitti_avg_error": self.computeSegmentAvgErr(seg)}
