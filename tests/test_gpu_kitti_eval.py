"""f-N4: KITTI sequence evaluation kernels (csrc/kitti_eval.cu through rslo_b200/utils/kitti_evaluation.py, the mirror
of the reference's kittiOdomEval / odom_to_abs_pose) against the numpy oracle (oracle/kitti_eval.py, pinned live against
the reference in tests/test_cpu_oracle.py).  float64 throughout; tolerances: chained poses 1e-9 absolute over 4541
frames, errors 1e-7 relative (+1e-11: arccos near 1), segment membership (which rows are valid, first frames, lengths) exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _odoms(n, seed, noise=0.0):
    rng = np.random.default_rng(seed)
    t = np.stack([rng.normal(1.0, 0.2, n), rng.normal(0, 0.03, n), rng.normal(0, 0.01, n)], 1)
    ang = rng.normal(0, 0.02, (n, 3))
    q = np.concatenate([np.ones((n, 1)), 0.5 * ang], 1)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    if noise:                                # prediction error in translation AND rotation (r_err = arccos(1 - eps) is
        r2 = np.random.default_rng(seed + 1)  # ill-conditioned at exactly equal rotations)
        t = t + r2.normal(0, noise, (n, 3))
        q = q + np.concatenate([np.zeros((n, 1)), r2.normal(0, 0.1 * noise, (n, 3))], 1)
        q /= np.linalg.norm(q, axis=1, keepdims=True)
    return np.concatenate([t, q], 1)


@pytest.mark.parametrize("n", [4541, 1101, 57])
def test_sequence_evaluation_matches_oracle(cuda, n):
    from oracle import kitti_eval as oke
    from rslo_b200.utils import kitti_evaluation as ke
    gts, preds = _odoms(n, 7), _odoms(n, 7, noise=0.02)
    a_gt, a_pr = oke.odom_to_abs_pose(gts), oke.odom_to_abs_pose(preds)
    d_pr, d_gt, dist = ke.odom_to_abs_pose_device(preds, gts)
    np.testing.assert_allclose(d_gt.cpu().numpy(), a_gt, rtol=0, atol=1e-9)
    np.testing.assert_allclose(d_pr.cpu().numpy(), a_pr, rtol=0, atol=1e-9)
    np.testing.assert_allclose(ke.odom_to_abs_pose(preds), a_pr, rtol=0, atol=1e-9)
    gt_mats = [oke.tq_to_RT(p) for p in a_gt]
    np.testing.assert_allclose(dist.cpu().numpy(), np.asarray(oke.trajectory_distances(gt_mats)), rtol=1e-12, atol=1e-9)
    ev = ke.kittiOdomEval()
    e_dev = ev.calcSequenceErrors(a_pr, a_gt)
    e_orc = oke.calc_sequence_errors(a_pr, a_gt)
    assert len(e_dev) == len(e_orc)
    if e_orc:
        d, o = np.asarray(e_dev), np.asarray(e_orc)
        assert np.array_equal(d[:, [0, 3]], o[:, [0, 3]])                 # same (first frame, length) rows, same order
        np.testing.assert_allclose(d, o, rtol=1e-7, atol=1e-11)
    s_dev, s_orc = ev.computeSegmentErr(e_dev), oke.segment_errors(e_orc)
    assert s_dev.keys() == s_orc.keys()
    np.testing.assert_allclose(ev.computeSegmentAvgErr(s_dev), oke.segment_avg(s_orc), rtol=1e-7)
    full = ev.evaluate_odometry(preds, gts)
    np.testing.assert_allclose(full["kitti_avg_error"], oke.segment_avg(s_orc), rtol=1e-7)
