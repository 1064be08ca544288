// K5: covariance-weighted point residual of the consistency loss, forward and backward (sm_100a).
//
// Replaces the per-pair body of Aleat5_1ChamferL2NormalWeightedALLSVDLoss._compute_loss
// (rslo/core/losses.py:348-363 span_cov2, :401-435 residual): ~120 eager ops, boolean-mask
// compactions (host syncs), batched 3x3 torch.inverse / torch.det via cuSOLVER.  Here one thread
// owns one associated point pair, everything 3x3 is closed form in registers, the ROI test
// (dist < threshold, losses.py:326-334) is a predicate instead of a compaction, and the sums are
// reduced per CTA and accumulated in double.
//
//   C(c)   = V diag(l1, l1+l2, l1+l2+l3) V^T,  V = rot(normalize(c[3:7] / (|c[3:7]| + 1e-9)))   (x,y,z,w)
//   Sigma  = C(cov_p[i]) + R C(cov_t[idx[i]]) R^T             (R = detached predicted rotation)
//   d      = p[i] - t[idx[i]]
//   loss   = mean_roi(d^T Sigma^-1 d) + reg * mean_roi(0.5 log det Sigma)
//
// Backward recomputes the per-point quantities (nothing but the inputs is saved) and scatters the
// gradients of the associated target rows with atomics.
#include "common.cuh"

namespace rslo {
namespace {

struct Sym3 {              // symmetric 3x3: xx xy xz yy yz zz
    float xx, xy, xz, yy, yz, zz;
};

__device__ __forceinline__ void quat_normalize(const float* c, float q[4], float& n1, float& n2)
{
    // losses.py:355 then kornia.quaternion_to_rotation_matrix's own F.normalize(eps=1e-12)
    n1 = sqrtf(c[3] * c[3] + c[4] * c[4] + c[5] * c[5] + c[6] * c[6]);
    const float s1 = 1.f / (n1 + 1e-9f);
    float a[4] = {c[3] * s1, c[4] * s1, c[5] * s1, c[6] * s1};
    n2 = sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + a[3] * a[3]);
    const float s2 = 1.f / fmaxf(n2, 1e-12f);
    q[0] = a[0] * s2; q[1] = a[1] * s2; q[2] = a[2] * s2; q[3] = a[3] * s2;
}

__device__ __forceinline__ void quat_to_rot(const float q[4], float V[3][3])
{
    const float x = q[0], y = q[1], z = q[2], w = q[3];
    const float tx = 2.f * x, ty = 2.f * y, tz = 2.f * z;
    const float twx = tx * w, twy = ty * w, twz = tz * w;
    const float txx = tx * x, txy = ty * x, txz = tz * x;
    const float tyy = ty * y, tyz = tz * y, tzz = tz * z;
    V[0][0] = 1.f - (tyy + tzz); V[0][1] = txy - twz;         V[0][2] = txz + twy;
    V[1][0] = txy + twz;         V[1][1] = 1.f - (txx + tzz); V[1][2] = tyz - twx;
    V[2][0] = txz - twy;         V[2][1] = tyz + twx;         V[2][2] = 1.f - (txx + tyy);
}

// C = V diag(lam) V^T
__device__ __forceinline__ Sym3 span_cov(const float* c, float V[3][3], float lam[3])
{
    float q[4], n1, n2;
    quat_normalize(c, q, n1, n2);
    quat_to_rot(q, V);
    lam[0] = c[0];
    lam[1] = lam[0] + c[1];
    lam[2] = lam[1] + c[2];
    Sym3 C;
    C.xx = V[0][0] * V[0][0] * lam[0] + V[0][1] * V[0][1] * lam[1] + V[0][2] * V[0][2] * lam[2];
    C.xy = V[0][0] * V[1][0] * lam[0] + V[0][1] * V[1][1] * lam[1] + V[0][2] * V[1][2] * lam[2];
    C.xz = V[0][0] * V[2][0] * lam[0] + V[0][1] * V[2][1] * lam[1] + V[0][2] * V[2][2] * lam[2];
    C.yy = V[1][0] * V[1][0] * lam[0] + V[1][1] * V[1][1] * lam[1] + V[1][2] * V[1][2] * lam[2];
    C.yz = V[1][0] * V[2][0] * lam[0] + V[1][1] * V[2][1] * lam[1] + V[1][2] * V[2][2] * lam[2];
    C.zz = V[2][0] * V[2][0] * lam[0] + V[2][1] * V[2][1] * lam[1] + V[2][2] * V[2][2] * lam[2];
    return C;
}

// R S R^T for symmetric S
__device__ __forceinline__ Sym3 rotate_sym(const float R[9], const Sym3& S)
{
    float M[3][3];     // M = R S
    const float s[3][3] = {{S.xx, S.xy, S.xz}, {S.xy, S.yy, S.yz}, {S.xz, S.yz, S.zz}};
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) M[i][j] = R[i * 3] * s[0][j] + R[i * 3 + 1] * s[1][j] + R[i * 3 + 2] * s[2][j];
    Sym3 o;
    o.xx = M[0][0] * R[0] + M[0][1] * R[1] + M[0][2] * R[2];
    o.xy = M[0][0] * R[3] + M[0][1] * R[4] + M[0][2] * R[5];
    o.xz = M[0][0] * R[6] + M[0][1] * R[7] + M[0][2] * R[8];
    o.yy = M[1][0] * R[3] + M[1][1] * R[4] + M[1][2] * R[5];
    o.yz = M[1][0] * R[6] + M[1][1] * R[7] + M[1][2] * R[8];
    o.zz = M[2][0] * R[6] + M[2][1] * R[7] + M[2][2] * R[8];
    return o;
}

// inverse and determinant of a symmetric 3x3 (adjugate form, as torch.inverse up to rounding)
__device__ __forceinline__ Sym3 inv_sym(const Sym3& S, float& det)
{
    const float A = S.yy * S.zz - S.yz * S.yz;
    const float B = S.xz * S.yz - S.xy * S.zz;
    const float C = S.xy * S.yz - S.xz * S.yy;
    det = S.xx * A + S.xy * B + S.xz * C;
    const float r = 1.f / det;
    Sym3 I;
    I.xx = A * r;
    I.xy = B * r;
    I.xz = C * r;
    I.yy = (S.xx * S.zz - S.xz * S.xz) * r;
    I.yz = (S.xy * S.xz - S.xx * S.yz) * r;
    I.zz = (S.xx * S.yy - S.xy * S.xy) * r;
    return I;
}

struct PointEval {
    Sym3 Sinv;
    float u[3];        // Sigma^-1 d
    float s, logdet;
};

__device__ __forceinline__ PointEval eval_point(const float* p, const float* t, const float* cp, const float* ct,
                                                const float R[9], float Vp[3][3], float lp[3], float Vt[3][3],
                                                float lt[3])
{
    Sym3 Cp = span_cov(cp, Vp, lp);
    Sym3 Ct = span_cov(ct, Vt, lt);
    Sym3 Cr = rotate_sym(R, Ct);
    Sym3 S{Cp.xx + Cr.xx, Cp.xy + Cr.xy, Cp.xz + Cr.xz, Cp.yy + Cr.yy, Cp.yz + Cr.yz, Cp.zz + Cr.zz};
    PointEval e;
    float det;
    e.Sinv = inv_sym(S, det);
    const float d0 = p[0] - t[0], d1 = p[1] - t[1], d2 = p[2] - t[2];
    e.u[0] = e.Sinv.xx * d0 + e.Sinv.xy * d1 + e.Sinv.xz * d2;
    e.u[1] = e.Sinv.xy * d0 + e.Sinv.yy * d1 + e.Sinv.yz * d2;
    e.u[2] = e.Sinv.xz * d0 + e.Sinv.yz * d1 + e.Sinv.zz * d2;
    e.s = d0 * e.u[0] + d1 * e.u[1] + d2 * e.u[2];
    e.logdet = 0.5f * logf(det);
    return e;
}

// sums[0] = sum s, sums[1] = sum 0.5 logdet, sums[2] = ROI count
__global__ void __launch_bounds__(256)
k_cov_fwd(const float* __restrict__ p, const float* __restrict__ t, const int* __restrict__ idx,
          const float* __restrict__ cov_p, const float* __restrict__ cov_t, const float* __restrict__ Rm,
          const float* __restrict__ dist, const float* __restrict__ thr, int n, double* __restrict__ sums)
{
    float R[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = Rm[i];
    const float th = *thr;
    float s = 0.f, l = 0.f, c = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (!(dist[i] < th)) continue;
        const int j = idx[i];
        float Vp[3][3], lp[3], Vt[3][3], lt[3];
        PointEval e = eval_point(p + i * 3, t + (size_t)j * 3, cov_p + (size_t)i * 7, cov_t + (size_t)j * 7, R, Vp, lp, Vt, lt);
        s += e.s;
        l += e.logdet;
        c += 1.f;
    }
    __shared__ float red[3][8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, d);
        l += __shfl_xor_sync(0xffffffffu, l, d);
        c += __shfl_xor_sync(0xffffffffu, c, d);
    }
    if (lane == 0) { red[0][wid] = s; red[1][wid] = l; red[2][wid] = c; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double v = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += red[threadIdx.x][w];
        atomicAdd(sums + threadIdx.x, v);
    }
}

__global__ void k_cov_finish(const double* __restrict__ sums, float reg, float* __restrict__ loss)
{
    if (threadIdx.x || blockIdx.x) return;
    const double cnt = sums[2];
    *loss = (float)(sums[0] / cnt + (double)reg * (sums[1] / cnt));     // cnt == 0 -> nan, as torch.mean of empty
}

// d(C)/d(c) for C = V diag(lam) V^T given the symmetric upstream G (dL/dC); writes 7 gradients.
__device__ __forceinline__ void span_cov_backward(const float* c, const float V[3][3], const float lam[3],
                                                  const float G[3][3], float gc[7])
{
    // d lam_j = v_j^T G v_j
    float dl[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float Gv0 = G[0][0] * V[0][j] + G[0][1] * V[1][j] + G[0][2] * V[2][j];
        float Gv1 = G[1][0] * V[0][j] + G[1][1] * V[1][j] + G[1][2] * V[2][j];
        float Gv2 = G[2][0] * V[0][j] + G[2][1] * V[1][j] + G[2][2] * V[2][j];
        dl[j] = V[0][j] * Gv0 + V[1][j] * Gv1 + V[2][j] * Gv2;
    }
    gc[0] = dl[0] + dl[1] + dl[2];
    gc[1] = dl[1] + dl[2];
    gc[2] = dl[2];
    // dV = 2 G V diag(lam)   (G symmetric)
    float dV[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            dV[i][j] = 2.f * lam[j] * (G[i][0] * V[0][j] + G[i][1] * V[1][j] + G[i][2] * V[2][j]);
    // rotation matrix -> unit quaternion (x,y,z,w)
    float q[4], n1, n2;
    quat_normalize(c, q, n1, n2);
    const float x = q[0], y = q[1], z = q[2], w = q[3];
    float dq[4];
    dq[0] = 2.f * (y * (dV[0][1] + dV[1][0]) + z * (dV[0][2] + dV[2][0]) - 2.f * x * (dV[1][1] + dV[2][2]) + w * (dV[2][1] - dV[1][2]));
    dq[1] = 2.f * (x * (dV[0][1] + dV[1][0]) + z * (dV[1][2] + dV[2][1]) - 2.f * y * (dV[0][0] + dV[2][2]) + w * (dV[0][2] - dV[2][0]));
    dq[2] = 2.f * (x * (dV[0][2] + dV[2][0]) + y * (dV[1][2] + dV[2][1]) - 2.f * z * (dV[0][0] + dV[1][1]) + w * (dV[1][0] - dV[0][1]));
    dq[3] = 2.f * (x * (dV[2][1] - dV[1][2]) + y * (dV[0][2] - dV[2][0]) + z * (dV[1][0] - dV[0][1]));
    // through q = a / max(|a|, 1e-12)
    const float m2 = fmaxf(n2, 1e-12f);
    float da[4];
    if (n2 >= 1e-12f) {
        const float dot = q[0] * dq[0] + q[1] * dq[1] + q[2] * dq[2] + q[3] * dq[3];
#pragma unroll
        for (int i = 0; i < 4; ++i) da[i] = (dq[i] - q[i] * dot) / m2;
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) da[i] = dq[i] / m2;
    }
    // through a = r / (|r| + 1e-9),  r = c[3:7]
    const float den = n1 + 1e-9f;
    const float rdot = c[3] * da[0] + c[4] * da[1] + c[5] * da[2] + c[6] * da[3];
    const float k = n1 > 0.f ? rdot / (n1 * den * den) : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) gc[3 + i] = da[i] / den - c[3 + i] * k;
}

__global__ void __launch_bounds__(128)
k_cov_bwd(const float* __restrict__ p, const float* __restrict__ t, const int* __restrict__ idx,
          const float* __restrict__ cov_p, const float* __restrict__ cov_t, const float* __restrict__ Rm,
          const float* __restrict__ dist, const float* __restrict__ thr, int n, float reg,
          const double* __restrict__ sums, const float* __restrict__ grad_loss, float* __restrict__ g_p,
          float* __restrict__ g_t, float* __restrict__ g_cov_p, float* __restrict__ g_cov_t)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float gc[7] = {0, 0, 0, 0, 0, 0, 0};
    float gp[3] = {0, 0, 0};
    if (dist[i] < *thr) {
        float R[9];
#pragma unroll
        for (int a = 0; a < 9; ++a) R[a] = Rm[a];
        const int j = idx[i];
        const float* cp = cov_p + (size_t)i * 7;
        const float* ct = cov_t + (size_t)j * 7;
        float Vp[3][3], lp[3], Vt[3][3], lt[3];
        PointEval e = eval_point(p + i * 3, t + (size_t)j * 3, cp, ct, R, Vp, lp, Vt, lt);
        const float gs = *grad_loss / (float)sums[2];
        // dL/dSigma = gs * (-u u^T + reg * 0.5 * Sigma^-1)
        const float h = 0.5f * reg;
        float G[3][3];
        G[0][0] = gs * (h * e.Sinv.xx - e.u[0] * e.u[0]);
        G[0][1] = gs * (h * e.Sinv.xy - e.u[0] * e.u[1]);
        G[0][2] = gs * (h * e.Sinv.xz - e.u[0] * e.u[2]);
        G[1][1] = gs * (h * e.Sinv.yy - e.u[1] * e.u[1]);
        G[1][2] = gs * (h * e.Sinv.yz - e.u[1] * e.u[2]);
        G[2][2] = gs * (h * e.Sinv.zz - e.u[2] * e.u[2]);
        G[1][0] = G[0][1]; G[2][0] = G[0][2]; G[2][1] = G[1][2];
        span_cov_backward(cp, Vp, lp, G, gc);
        // target covariance sees R^T G R
        float M[3][3], Gt[3][3];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) M[a][b] = R[0 * 3 + a] * G[0][b] + R[1 * 3 + a] * G[1][b] + R[2 * 3 + a] * G[2][b];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) Gt[a][b] = M[a][0] * R[0 * 3 + b] + M[a][1] * R[1 * 3 + b] + M[a][2] * R[2 * 3 + b];
        float gt7[7];
        span_cov_backward(ct, Vt, lt, Gt, gt7);
#pragma unroll
        for (int a = 0; a < 7; ++a) atomicAdd(g_cov_t + (size_t)j * 7 + a, gt7[a]);
        // d = p - t: dL/dp = 2 gs u, dL/dt = -2 gs u
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            gp[a] = 2.f * gs * e.u[a];
            atomicAdd(g_t + (size_t)j * 3 + a, -gp[a]);
        }
    }
#pragma unroll
    for (int a = 0; a < 7; ++a) g_cov_p[(size_t)i * 7 + a] = gc[a];
    if (g_p) {
#pragma unroll
        for (int a = 0; a < 3; ++a) g_p[i * 3 + a] = gp[a];
    }
}

}  // namespace
}  // namespace rslo

using namespace rslo;

extern "C" int rslo_cov_residual_forward(const float* pred, const float* target, const int32_t* idx,
                                         const float* cov_pred, const float* cov_target, const float* R,
                                         const float* dist, const float* dist_threshold, int n, float reg_weight,
                                         double* sums, float* loss, rslo_stream_t stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    RSLO_CHECK(cudaMemsetAsync(sums, 0, 4 * sizeof(double), st));
    if (n > 0) {
        int blocks = cdiv(n, 256);
        if (blocks > 148 * 4) blocks = 148 * 4;
        RSLO_COUNT();
        k_cov_fwd<<<blocks, 256, 0, st>>>(pred, target, idx, cov_pred, cov_target, R, dist, dist_threshold, n, sums);
    }
    RSLO_COUNT();
    k_cov_finish<<<1, 32, 0, st>>>(sums, reg_weight, loss);
    RSLO_CHECK_LAUNCH("rslo_cov_residual_forward");
    return 0;
}

extern "C" int rslo_cov_residual_backward(const float* pred, const float* target, const int32_t* idx,
                                          const float* cov_pred, const float* cov_target, const float* R,
                                          const float* dist, const float* dist_threshold, int n, int m,
                                          float reg_weight, const double* sums, const float* grad_loss,
                                          float* grad_pred, float* grad_target, float* grad_cov_pred,
                                          float* grad_cov_target, rslo_stream_t stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    RSLO_CHECK(cudaMemsetAsync(grad_target, 0, (size_t)m * 3 * sizeof(float), st));
    RSLO_CHECK(cudaMemsetAsync(grad_cov_target, 0, (size_t)m * 7 * sizeof(float), st));
    if (n > 0) {
        RSLO_COUNT();
        k_cov_bwd<<<cdiv(n, 128), 128, 0, st>>>(pred, target, idx, cov_pred, cov_target, R, dist, dist_threshold, n,
                                                reg_weight, sums, grad_loss, grad_pred, grad_target, grad_cov_pred,
                                                grad_cov_target);
    }
    RSLO_CHECK_LAUNCH("rslo_cov_residual_backward");
    return 0;
}
