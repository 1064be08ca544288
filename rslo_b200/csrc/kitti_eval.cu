// KITTI odometry sequence evaluation on the device (SURVEY §8 row f-N4), float64.
//
// Replaces the pure-Python loops that run on rank 0 while every GPU waits at a barrier (`train_hdf5.py:872-886`):
//   geometric.odom_to_abs_pose          rslo/utils/geometric.py:376-406   (chain of relative poses -> absolute poses)
//   kittiOdomEval.trajectoryDistances   rslo/utils/kitti_evaluation.py:42-61
//   kittiOdomEval.calcSequenceErrors    rslo/utils/kitti_evaluation.py:95-145 (every 10th start frame x 8 segment lengths)
//   kittiOdomEval.computeSegmentErr / computeSegmentAvgErr                  :157-198
// The pose chain is sequential by definition (each step renormalises the quaternion with q / (|q| + 1e-6), which is not
// associative), so one thread walks it - 4541 frames of KITTI-00 take ~1 ms instead of the reference's ~0.3 s of numpy
// calls; predictions and ground truth are walked by two CTAs at once.  The segment errors are one thread per
// (start frame, length): the end frame by binary search over the cumulative distances (the reference scans linearly;
// the distances are non-decreasing, so the first index above the threshold is the same), then 3x4 pose algebra.
#include "common.cuh"

namespace rslo {
namespace {

struct Q4 {
    double w, x, y, z;
};

// pose_utils_np.qmult + normalize(eps 1e-6)  (rslo/utils/pose_utils_np.py:127-163)
__device__ Q4 qmult_norm(const Q4& a, const Q4& b)
{
    Q4 r;
    r.w = a.w * b.w - ((a.x * b.x + a.y * b.y) + a.z * b.z);
    r.x = (a.x * b.w + b.x * a.w) + (a.y * b.z - a.z * b.y);
    r.y = (a.y * b.w + b.y * a.w) + (a.z * b.x - a.x * b.z);
    r.z = (a.z * b.w + b.z * a.w) + (a.x * b.y - a.y * b.x);
    const double n = sqrt(((r.w * r.w + r.x * r.x) + r.y * r.y) + r.z * r.z) + 1e-6;
    r.w /= n; r.x /= n; r.y /= n; r.z /= n;
    return r;
}
// pose_utils_np.rotate_vec_by_q: t + 2 qs (qv x t) + 2 qv x (qv x t)   (:228-243)
__device__ void rotate_by_q(const double* t, const Q4& q, double* o)
{
    const double bx = q.y * t[2] - q.z * t[1], by = q.z * t[0] - q.x * t[2], bz = q.x * t[1] - q.y * t[0];
    const double cx = 2.0 * (q.y * bz - q.z * by), cy = 2.0 * (q.z * bx - q.x * bz), cz = 2.0 * (q.x * by - q.y * bx);
    o[0] = (t[0] + 2.0 * bx * q.w) + cx;
    o[1] = (t[1] + 2.0 * by * q.w) + cy;
    o[2] = (t[2] + 2.0 * bz * q.w) + cz;
}

// blockIdx.x = sequence (0: predictions, 1: ground truth).  abs[0] = identity; the running pose starts at odoms[0]
// (geometric.py:388-405).  The ground-truth CTA also accumulates the trajectory distances.
__global__ void k_abs_pose(const double* __restrict__ odom_a, const double* __restrict__ odom_b, int n,
                           double* __restrict__ abs_a, double* __restrict__ abs_b, double* __restrict__ dist_b)
{
    if (threadIdx.x != 0 || n <= 0) return;
    const double* od = blockIdx.x == 0 ? odom_a : odom_b;
    double* ab = blockIdx.x == 0 ? abs_a : abs_b;
    if (od == nullptr || ab == nullptr) return;
    ab[0] = ab[1] = ab[2] = 0.0; ab[3] = 1.0; ab[4] = ab[5] = ab[6] = 0.0;
    double tp[3] = {od[0], od[1], od[2]};
    Q4 rp = {od[3], od[4], od[5], od[6]};
    for (int i = 1; i < n; ++i) {
        const double* o = od + (size_t)i * 7;
        const Q4 rc = {o[3], o[4], o[5], o[6]};
        double rot[3];
        rotate_by_q(o, rp, rot);
        tp[0] += rot[0]; tp[1] += rot[1]; tp[2] += rot[2];
        rp = qmult_norm(rp, rc);
        double* d = ab + (size_t)i * 7;
        d[0] = tp[0]; d[1] = tp[1]; d[2] = tp[2]; d[3] = rp.w; d[4] = rp.x; d[5] = rp.y; d[6] = rp.z;
    }
    if (blockIdx.x == 1 && dist_b != nullptr) {
        double acc = 0.0;
        dist_b[0] = 0.0;
        for (int i = 0; i + 1 < n; ++i) {
            const double dx = ab[i * 7] - ab[(i + 1) * 7], dy = ab[i * 7 + 1] - ab[(i + 1) * 7 + 1], dz = ab[i * 7 + 2] - ab[(i + 1) * 7 + 2];
            acc += sqrt((dx * dx + dy * dy) + dz * dz);
            dist_b[i + 1] = acc;
        }
    }
}

// geometric.tq_to_RT: rotation of the (w,x,y,z) quaternion as numpy-quaternion's as_rotation_matrix gives it
// (general, non-unit form: entries scaled by 2 / |q|^2)
__device__ void tq_to_RT(const double* tq, double* R, double* t)
{
    const double w = tq[3], x = tq[4], y = tq[5], z = tq[6];
    const double s = 2.0 / (((w * w + x * x) + y * y) + z * z);
    R[0] = 1.0 - s * (y * y + z * z); R[1] = s * (x * y - z * w); R[2] = s * (x * z + y * w);
    R[3] = s * (x * y + z * w); R[4] = 1.0 - s * (x * x + z * z); R[5] = s * (y * z - x * w);
    R[6] = s * (x * z - y * w); R[7] = s * (y * z + x * w); R[8] = 1.0 - s * (x * x + y * y);
    t[0] = tq[0]; t[1] = tq[1]; t[2] = tq[2];
}
// D = inv(A) B for rigid [R|t]
__device__ void rel(const double* Ra, const double* ta, const double* Rb, const double* tb, double* R, double* t)
{
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[i * 3 + j] = (Ra[i] * Rb[j] + Ra[3 + i] * Rb[3 + j]) + Ra[6 + i] * Rb[6 + j];
    const double d[3] = {tb[0] - ta[0], tb[1] - ta[1], tb[2] - ta[2]};
    for (int i = 0; i < 3; ++i) t[i] = (Ra[i] * d[0] + Ra[3 + i] * d[1]) + Ra[6 + i] * d[2];
}

// one thread per (start frame, segment length): err[row] = {first, r_err/len, t_err/len, len, speed}, valid[row]
__global__ void k_seq_errors(const double* __restrict__ abs_pred, int n_pred, const double* __restrict__ abs_gt, int n_gt,
                             const double* __restrict__ dist, int step, double* __restrict__ err, int* __restrict__ valid,
                             int rows)
{
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    const int first = (row >> 3) * step;
    const double len = 100.0 * ((row & 7) + 1);
    // first i >= first with dist[i] > dist[first] + len
    const double thr = dist[first] + len;
    int lo = first, hi = n_gt;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (dist[mid] > thr) hi = mid;
        else lo = mid + 1;
    }
    const int last = lo < n_gt ? lo : -1;
    double* e = err + (size_t)row * 5;
    if (last == -1 || last >= n_pred || first >= n_pred) {
        valid[row] = 0;
        e[0] = e[1] = e[2] = e[3] = e[4] = 0.0;
        return;
    }
    double Rf[9], tf[3], Rl[9], tl[3], Rg[9], tg[3], Rr[9], tr[3], Re[9], te[3];
    tq_to_RT(abs_gt + (size_t)first * 7, Rf, tf);
    tq_to_RT(abs_gt + (size_t)last * 7, Rl, tl);
    rel(Rf, tf, Rl, tl, Rg, tg);                       // pose_delta_gt
    tq_to_RT(abs_pred + (size_t)first * 7, Rf, tf);
    tq_to_RT(abs_pred + (size_t)last * 7, Rl, tl);
    rel(Rf, tf, Rl, tl, Rr, tr);                       // pose_delta_result
    rel(Rr, tr, Rg, tg, Re, te);                       // pose_error = inv(result) gt
    const double d = 0.5 * (((Re[0] + Re[4]) + Re[8]) - 1.0);
    const double r_err = acos(fmax(fmin(d, 1.0), -1.0));
    const double t_err = sqrt((te[0] * te[0] + te[1] * te[1]) + te[2] * te[2]);
    const double num_frames = (double)(last - first) + 1.0;
    valid[row] = 1;
    e[0] = (double)first; e[1] = r_err / len; e[2] = t_err / len; e[3] = len; e[4] = len / (0.1 * num_frames);
}

}  // namespace
}  // namespace rslo

using namespace rslo;

extern "C" int rslo_odom_to_abs_pose(const double* odom_a, const double* odom_b, int n, double* abs_a, double* abs_b,
                                     double* dist_b, rslo_stream_t stream)
{
    if (n <= 0) return 0;
    RSLO_COUNT();
    k_abs_pose<<<2, 32, 0, (cudaStream_t)stream>>>(odom_a, odom_b, n, abs_a, abs_b, dist_b);
    RSLO_CHECK_LAUNCH("rslo_odom_to_abs_pose");
    return 0;
}

extern "C" int rslo_kitti_sequence_errors(const double* abs_pred, int n_pred, const double* abs_gt, int n_gt,
                                          const double* dist_gt, int step, double* err, int32_t* valid, rslo_stream_t stream)
{
    if (n_gt <= 0 || step <= 0) return 0;
    const int rows = ((n_gt + step - 1) / step) * 8;
    RSLO_COUNT();
    k_seq_errors<<<cdiv(rows, 128), 128, 0, (cudaStream_t)stream>>>(abs_pred, n_pred, abs_gt, n_gt, dist_gt, step, err, valid, rows);
    RSLO_CHECK_LAUNCH("rslo_kitti_sequence_errors");
    return 0;
}
