// Predicted pose applied to the target frame's points: y = R(q) x + t, forward and backward (sm_100a).
//
// Replaces the torch glue of `rslo/models/voxel_odom_net.py:671-690`: kornia.quaternion_to_rotation_matrix (0.4.0:
// L2-normalise with eps 1e-12, then the 9 quadratic entries; ~30 elementwise launches with as many autograd nodes),
// `p1[:, :, :3] @ R^T + T` (batched GEMM + add) and, backward, the same chain reversed - per frame pair and step.
// One streaming launch forward (also writes R for the covariance rotation of the loss), one reduction launch
// backward: dt = sum g, dR = sum g x^T in double, then the Jacobian of R(q) and of the normalisation in the last CTA.
#include "common.cuh"

namespace rslo {
namespace {

constexpr int PX_THREADS = 256;

// kornia 0.4.0 quaternion_to_rotation_matrix on (x,y,z,w); q here is (w,x,y,z)
__device__ __forceinline__ void quat_to_R(const float* __restrict__ q, float* R, float* qn, float* norm)
{
    const float n = fmaxf(sqrtf(((q[1] * q[1] + q[2] * q[2]) + q[3] * q[3]) + q[0] * q[0]), 1e-12f);
    const float x = q[1] / n, y = q[2] / n, z = q[3] / n, w = q[0] / n;
    const float tx = 2.f * x, ty = 2.f * y, tz = 2.f * z;
    const float twx = tx * w, twy = ty * w, twz = tz * w;
    const float txx = tx * x, txy = ty * x, txz = tz * x;
    const float tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1.f - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1.f - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1.f - (txx + tyy);
    if (qn) {
        qn[0] = w; qn[1] = x; qn[2] = y; qn[3] = z;
        *norm = n;
    }
}

__global__ void __launch_bounds__(PX_THREADS)
k_pair_transform_fwd(const float* __restrict__ x, int ldx, int n, const float* __restrict__ q, const float* __restrict__ t,
                     int identity, float* __restrict__ y, float* __restrict__ R_out)
{
    float R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, T[3] = {0, 0, 0};
    if (!identity) {
        quat_to_R(q, R, nullptr, nullptr);
        T[0] = t[0]; T[1] = t[1]; T[2] = t[2];
    }
    if (blockIdx.x == 0 && threadIdx.x < 9) R_out[threadIdx.x] = R[threadIdx.x];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float a = __ldg(x + (size_t)i * ldx), b = __ldg(x + (size_t)i * ldx + 1), c = __ldg(x + (size_t)i * ldx + 2);
        // (x @ R^T + t): row i of the batched GEMM the reference runs, accumulated in its order
        y[(size_t)i * 3 + 0] = ((a * R[0] + b * R[1]) + c * R[2]) + T[0];
        y[(size_t)i * 3 + 1] = ((a * R[3] + b * R[4]) + c * R[5]) + T[1];
        y[(size_t)i * 3 + 2] = ((a * R[6] + b * R[7]) + c * R[8]) + T[2];
    }
}

// acc[0..8] = sum_i g_i x_i^T (row-major), acc[9..11] = sum_i g_i; last CTA: dq (w,x,y,z), dt
__global__ void __launch_bounds__(PX_THREADS)
k_pair_transform_bwd(const float* __restrict__ g, const float* __restrict__ x, int ldx, int n, const float* __restrict__ q,
                     double* __restrict__ acc, unsigned int* __restrict__ counter, float* __restrict__ dq,
                     float* __restrict__ dt)
{
    __shared__ double sh[12][PX_THREADS / 32];
    __shared__ bool last;
    double s[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) s[k] = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float gx = __ldg(g + (size_t)i * 3), gy = __ldg(g + (size_t)i * 3 + 1), gz = __ldg(g + (size_t)i * 3 + 2);
        const float a = __ldg(x + (size_t)i * ldx), b = __ldg(x + (size_t)i * ldx + 1), c = __ldg(x + (size_t)i * ldx + 2);
        s[0] += (double)(gx * a); s[1] += (double)(gx * b); s[2] += (double)(gx * c);
        s[3] += (double)(gy * a); s[4] += (double)(gy * b); s[5] += (double)(gy * c);
        s[6] += (double)(gz * a); s[7] += (double)(gz * b); s[8] += (double)(gz * c);
        s[9] += (double)gx; s[10] += (double)gy; s[11] += (double)gz;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 12; ++k) {
        double v = s[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sh[k][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        double v = 0.0;
        for (int w = 0; w < PX_THREADS / 32; ++w) v += sh[threadIdx.x][w];
        atomicAdd(acc + threadIdx.x, v);
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        double G[12];
        for (int k = 0; k < 12; ++k) {
            G[k] = ((volatile double*)acc)[k];
            acc[k] = 0.0;                              // scratch left zeroed for the next call
        }
        *counter = 0;
        dt[0] = (float)G[9]; dt[1] = (float)G[10]; dt[2] = (float)G[11];
        float R[9], qn[4], norm;
        quat_to_R(q, R, qn, &norm);
        const double w = qn[0], xq = qn[1], yq = qn[2], zq = qn[3];
        // R00 = 1-2y^2-2z^2  R01 = 2xy-2wz  R02 = 2xz+2wy  R10 = 2xy+2wz  R11 = 1-2x^2-2z^2  R12 = 2yz-2wx
        // R20 = 2xz-2wy  R21 = 2yz+2wx  R22 = 1-2x^2-2y^2
        const double dx = 2.0 * (yq * (G[1] + G[3]) + zq * (G[2] + G[6]) + w * (G[7] - G[5])) - 4.0 * xq * (G[4] + G[8]);
        const double dy = 2.0 * (xq * (G[1] + G[3]) + zq * (G[5] + G[7]) + w * (G[2] - G[6])) - 4.0 * yq * (G[0] + G[8]);
        const double dz = 2.0 * (xq * (G[2] + G[6]) + yq * (G[5] + G[7]) + w * (G[3] - G[1])) - 4.0 * zq * (G[0] + G[4]);
        const double dw = 2.0 * (zq * (G[3] - G[1]) + yq * (G[2] - G[6]) + xq * (G[7] - G[5]));
        // through qn = q / max(|q|, eps)
        const double raw = sqrt((double)q[0] * q[0] + (double)q[1] * q[1] + (double)q[2] * q[2] + (double)q[3] * q[3]);
        double o[4] = {dw, dx, dy, dz};
        if (raw > 1e-12) {
            const double dot = w * dw + xq * dx + yq * dy + zq * dz;
            o[0] = (dw - w * dot) / norm; o[1] = (dx - xq * dot) / norm; o[2] = (dy - yq * dot) / norm; o[3] = (dz - zq * dot) / norm;
        } else {
            for (int k = 0; k < 4; ++k) o[k] /= norm;
        }
        for (int k = 0; k < 4; ++k) dq[k] = (float)o[k];
    }
}

}  // namespace
}  // namespace rslo

using namespace rslo;

extern "C" size_t rslo_pair_transform_workspace_bytes(void) { return 12 * sizeof(double) + 64; }

extern "C" int rslo_pair_transform_forward(const float* x, int ldx, int n, const float* q_wxyz, const float* t, int identity,
                                           float* y, float* R_out, rslo_stream_t stream)
{
    int ctas = cdiv(n, PX_THREADS);
    if (ctas < 1) ctas = 1;
    if (ctas > 148 * 4) ctas = 148 * 4;
    RSLO_COUNT();
    k_pair_transform_fwd<<<ctas, PX_THREADS, 0, (cudaStream_t)stream>>>(x, ldx, n, q_wxyz, t, identity, y, R_out);
    RSLO_CHECK_LAUNCH("rslo_pair_transform_forward");
    return 0;
}

extern "C" int rslo_pair_transform_backward(const float* grad_y, const float* x, int ldx, int n, const float* q_wxyz,
                                            float* dq_wxyz, float* dt, void* workspace, size_t workspace_bytes,
                                            rslo_stream_t stream)
{
    if (workspace == nullptr || workspace_bytes < rslo_pair_transform_workspace_bytes()) {
        set_last_error("rslo_pair_transform_backward: workspace too small", cudaErrorInvalidValue);
        return (int)cudaErrorInvalidValue;
    }
    double* acc = (double*)workspace;                                   // zeroed once by the caller, left zeroed
    unsigned int* counter = (unsigned int*)((char*)workspace + 12 * sizeof(double));
    int ctas = cdiv(n, PX_THREADS * 4);
    if (ctas < 1) ctas = 1;
    if (ctas > 148) ctas = 148;
    RSLO_COUNT();
    k_pair_transform_bwd<<<ctas, PX_THREADS, 0, (cudaStream_t)stream>>>(grad_y, x, ldx, n, q_wxyz, acc, counter, dq_wxyz, dt);
    RSLO_CHECK_LAUNCH("rslo_pair_transform_backward");
    return 0;
}
