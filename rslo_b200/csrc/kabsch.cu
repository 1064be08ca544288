// K6: weighted Kabsch alignment without host round trips (sm_100a).
//
// Replaces SVDHead.forward (rslo/layers/svd.py:13-64) as used by the consistency loss's ICP
// refinement (rslo/core/losses.py:440-456): there every call costs a boolean-mask compaction
// (nonzero -> host sync), a 3x3 torch.svd (cuSOLVER) and a `det < 0` branch on the host.  Here the
// ROI mask is applied as a 0/1 factor inside one reduction over all points and the 3x3 SVD is a
// Jacobi eigen-solve on one thread, so the whole refinement stays on the stream.
//   means are UNWEIGHTED over the masked points (svd.py:27-28)
//   H = sum_i m_i w_i (x_i - xbar)(y_i - ybar)^T         (svd.py:33)
//   R = V U^T, last column of V negated if det < 0        (svd.py:36-44)
//   t = -R xbar + ybar; returns R^T and -R^T t            (svd.py:57-64)
// Raw moments are accumulated in double, so the one-pass form loses nothing against the reference's
// centred fp32 form; inputs are detached in the reference, so there is no backward.
#include "common.cuh"

namespace rslo {
namespace {

constexpr int KB_NACC = 26;  // m, mx(3), my(3), mw, mwx(3), mwy(3), mwxy(9), pad -> 26 used: 1+3+3+1+3+3+9 = 23

__global__ void __launch_bounds__(256)
k_kabsch_accum(const float* __restrict__ src, const float* __restrict__ tgt, const int* __restrict__ tgt_idx,
               const float* __restrict__ weight, const float* __restrict__ normal, const float* __restrict__ mask,
               const float* __restrict__ dist, const float* __restrict__ thr, int n, double* __restrict__ acc)
{
    double a[23];
#pragma unroll
    for (int i = 0; i < 23; ++i) a[i] = 0.0;
    const float th = thr ? *thr : 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float m = mask ? mask[i] : 1.f;
        if (dist) m = dist[i] < th ? m : 0.f;          // ROI: dist < threshold (losses.py:331)
        if (m == 0.f) continue;
        const int ti = tgt_idx ? tgt_idx[i] : i;                 // association gather fused (losses.py:405,476)
        const float xf[3] = {src[i * 3], src[i * 3 + 1], src[i * 3 + 2]};
        const float yf[3] = {tgt[(size_t)ti * 3], tgt[(size_t)ti * 3 + 1], tgt[(size_t)ti * 3 + 2]};
        double w = weight ? (double)weight[i] : 1.0;
        if (normal) {
            // w = |cos(normal_i, y - x)|^2 (losses.py:411,440-456: cosine_similarity(...).abs(), squared at the call)
            const float nx = normal[i * 3], ny = normal[i * 3 + 1], nz = normal[i * 3 + 2];
            const float vx = yf[0] - xf[0], vy = yf[1] - xf[1], vz = yf[2] - xf[2];
            const float nn = fmaxf(sqrtf(nx * nx + ny * ny + nz * nz), 1e-8f);
            const float vn = fmaxf(sqrtf(vx * vx + vy * vy + vz * vz), 1e-8f);
            const float c = (nx * vx + ny * vy + nz * vz) / (nn * vn);
            w *= (double)(c * c);
        }
        const double x[3] = {xf[0], xf[1], xf[2]};
        const double y[3] = {yf[0], yf[1], yf[2]};
        const double dm = m, mw = dm * w;
        a[0] += dm;
        a[7] += mw;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            a[1 + r] += dm * x[r];
            a[4 + r] += dm * y[r];
            a[8 + r] += mw * x[r];
            a[11 + r] += mw * y[r];
#pragma unroll
            for (int c = 0; c < 3; ++c) a[14 + r * 3 + c] += mw * x[r] * y[c];
        }
    }
    __shared__ double red[8][23];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 23; ++i) {
        double v = a[i];
#pragma unroll
        for (int s = 16; s; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
        if (lane == 0) red[wid][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < 23) {
        double v = 0;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) v += red[w8][threadIdx.x];
        atomicAdd(acc + threadIdx.x, v);
    }
}

// Jacobi eigen-decomposition of a symmetric 3x3 (double). M is destroyed; V columns = eigenvectors.
__device__ void jacobi3(double M[3][3], double V[3][3], double ev[3])
{
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) V[i][j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = fabs(M[0][1]) + fabs(M[0][2]) + fabs(M[1][2]);
        double diag = fabs(M[0][0]) + fabs(M[1][1]) + fabs(M[2][2]);
        if (off <= 1e-300 || off <= 1e-17 * diag) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (M[p][q] == 0.0) continue;
                double theta = (M[q][q] - M[p][p]) / (2.0 * M[p][q]);
                double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {        // M <- M J
                    double mkp = M[k][p], mkq = M[k][q];
                    M[k][p] = c * mkp - s * mkq;
                    M[k][q] = s * mkp + c * mkq;
                }
                for (int k = 0; k < 3; ++k) {        // M <- J^T M
                    double mpk = M[p][k], mqk = M[q][k];
                    M[p][k] = c * mpk - s * mqk;
                    M[q][k] = s * mpk + c * mqk;
                }
                for (int k = 0; k < 3; ++k) {
                    double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq;
                    V[k][q] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < 3; ++i) ev[i] = M[i][i];
}

__device__ __forceinline__ void cross3(const double a[3], const double b[3], double c[3])
{
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

// One thread: means, H, SVD, R, t.  Optionally composes the running ICP transform
// (res_r <- R res_r, res_t <- R res_t + t; losses.py:463-465) when comp_R/comp_t are given.
__global__ void k_kabsch_solve(const double* __restrict__ acc, float* __restrict__ R_out, float* __restrict__ t_out,
                               float* __restrict__ comp_R, float* __restrict__ comp_t)
{
    if (threadIdx.x || blockIdx.x) return;
    const double m = acc[0];
    double Rf[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, tf[3] = {0, 0, 0};
    if (m > 0.0) {
        double xb[3], yb[3], H[3][3];
        for (int r = 0; r < 3; ++r) { xb[r] = acc[1 + r] / m; yb[r] = acc[4 + r] / m; }
        const double sw = acc[7];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
                H[r][c] = acc[14 + r * 3 + c] - xb[r] * acc[11 + c] - acc[8 + r] * yb[c] + sw * xb[r] * yb[c];
        // H = U S V^T ; eigen-decompose H^T H = V S^2 V^T
        double M[3][3], V[3][3], ev[3];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double s = 0;
                for (int k = 0; k < 3; ++k) s += H[k][i] * H[k][j];
                M[i][j] = s;
            }
        jacobi3(M, V, ev);
        int ord[3] = {0, 1, 2};                      // descending singular values, as torch.svd
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2 - i; ++j)
                if (ev[ord[j]] < ev[ord[j + 1]]) { int tmp = ord[j]; ord[j] = ord[j + 1]; ord[j + 1] = tmp; }
        double Vs[3][3], U[3][3];                    // column c of Vs = eigenvector ord[c]
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) Vs[r][c] = V[r][ord[c]];
        double u[3][3];                              // u[c] = column c of U
        for (int c = 0; c < 3; ++c) {
            for (int r = 0; r < 3; ++r) u[c][r] = H[r][0] * Vs[0][c] + H[r][1] * Vs[1][c] + H[r][2] * Vs[2][c];
        }
        double n0 = sqrt(u[0][0] * u[0][0] + u[0][1] * u[0][1] + u[0][2] * u[0][2]);
        bool ok = n0 > 0.0;
        if (ok) {
            for (int r = 0; r < 3; ++r) u[0][r] /= n0;
            // Gram-Schmidt the second column for robustness when s2 is small
            double d01 = u[0][0] * u[1][0] + u[0][1] * u[1][1] + u[0][2] * u[1][2];
            for (int r = 0; r < 3; ++r) u[1][r] -= d01 * u[0][r];
            double n1 = sqrt(u[1][0] * u[1][0] + u[1][1] * u[1][1] + u[1][2] * u[1][2]);
            ok = n1 > 1e-300;
            if (ok) {
                for (int r = 0; r < 3; ++r) u[1][r] /= n1;
                double c3[3];
                cross3(u[0], u[1], c3);
                double sgn = c3[0] * u[2][0] + c3[1] * u[2][1] + c3[2] * u[2][2];
                double s3 = sgn < 0.0 ? -1.0 : 1.0;
                for (int r = 0; r < 3; ++r) u[2][r] = s3 * c3[r];
            }
        }
        if (ok) {
            for (int c = 0; c < 3; ++c)
                for (int r = 0; r < 3; ++r) U[r][c] = u[c][r];
            double R[3][3];
            for (int pass = 0; pass < 2; ++pass) {
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) R[i][j] = Vs[i][0] * U[j][0] + Vs[i][1] * U[j][1] + Vs[i][2] * U[j][2];
                double det = R[0][0] * (R[1][1] * R[2][2] - R[1][2] * R[2][1]) -
                             R[0][1] * (R[1][0] * R[2][2] - R[1][2] * R[2][0]) +
                             R[0][2] * (R[1][0] * R[2][1] - R[1][1] * R[2][0]);
                if (det >= 0.0 || pass == 1) break;
                for (int r = 0; r < 3; ++r) Vs[r][2] = -Vs[r][2];      // v <- v @ reflect (svd.py:42)
            }
            // t = -R xbar + ybar ; return R^T, -R^T t
            double t[3];
            for (int i = 0; i < 3; ++i) t[i] = -(R[i][0] * xb[0] + R[i][1] * xb[1] + R[i][2] * xb[2]) + yb[i];
            for (int i = 0; i < 3; ++i) {
                for (int j = 0; j < 3; ++j) Rf[i][j] = R[j][i];
            }
            for (int i = 0; i < 3; ++i) tf[i] = -(Rf[i][0] * t[0] + Rf[i][1] * t[1] + Rf[i][2] * t[2]);
        }
    }
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) R_out[i * 3 + j] = (float)Rf[i][j];
        t_out[i] = (float)tf[i];
    }
    if (comp_R) {
        float Rn[9], tn[3];
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j)
                Rn[i * 3 + j] = (float)Rf[i][0] * comp_R[0 * 3 + j] + (float)Rf[i][1] * comp_R[1 * 3 + j] +
                                (float)Rf[i][2] * comp_R[2 * 3 + j];
            tn[i] = (float)Rf[i][0] * comp_t[0] + (float)Rf[i][1] * comp_t[1] + (float)Rf[i][2] * comp_t[2] +
                    (float)tf[i];
        }
        for (int i = 0; i < 9; ++i) comp_R[i] = Rn[i];
        for (int i = 0; i < 3; ++i) comp_t[i] = tn[i];
    }
}

}  // namespace
}  // namespace rslo

using namespace rslo;

extern "C" size_t rslo_kabsch_workspace_bytes(void) { return 32 * sizeof(double); }

extern "C" int rslo_kabsch(const float* src, const float* tgt, const int32_t* tgt_idx, const float* weight,
                           const float* normal, const float* mask, const float* dist, const float* dist_threshold,
                           int n, float* R_out, float* t_out, float* comp_R, float* comp_t, void* workspace,
                           size_t workspace_bytes, rslo_stream_t stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (workspace_bytes < 32 * sizeof(double)) {
        set_last_error("rslo_kabsch: workspace too small", cudaErrorMemoryAllocation);
        return (int)cudaErrorMemoryAllocation;
    }
    double* acc = (double*)workspace;
    RSLO_CHECK(cudaMemsetAsync(acc, 0, 32 * sizeof(double), st));
    if (n > 0) {
        int blocks = cdiv(n, 256);
        if (blocks > 148 * 2) blocks = 148 * 2;
        RSLO_COUNT();
        k_kabsch_accum<<<blocks, 256, 0, st>>>(src, tgt, tgt_idx, weight, normal, mask, dist, dist_threshold, n, acc);
    }
    RSLO_COUNT();
    k_kabsch_solve<<<1, 32, 0, st>>>(acc, R_out, t_out, comp_R, comp_t);
    RSLO_CHECK_LAUNCH("rslo_kabsch");
    return 0;
}
