// BatchNorm1d (+ LeakyReLU) over the rows of a sparse level, one statistics group per frame (sm_100a).
//
// Replaces `nn.BatchNorm1d(C)` + `nn.LeakyReLU` after the five convolutions of the covariance decoder
// (`rslo/models/middle.py:181-213`, applied by spconv.SparseSequential to `.features`).  The reference runs the
// encoder once per frame, so each frame is normalised with its own batch statistics and the running statistics
// move once per frame; here the T frames of a step share one pass and their rows are stacked, so the kernels take
// the frames' row ranges ("segments") and keep statistics per segment.  torch's native_batch_norm needs 3 launches
// per frame forward and 3 backward (+ a LeakyReLU pass each way, + `grad += ` per frame for gamma / beta):
// ~120 launches per step for 5 layers x 4 frames; this is 2 launches per layer forward, 2 backward.
//
// Streaming kernels (HBM/L2-bound): forward reads x twice (statistics, apply) and writes z; backward reads dz, x
// twice and writes dx.  Statistics are reduced per CTA in fp32 over <= 32 rows per thread and across CTAs in double.
#include "common.cuh"

namespace rslo {
namespace {

constexpr int SB_THREADS = 256;
constexpr int SB_MAX_SEG = 16;

struct SegTable {
    int off[SB_MAX_SEG + 1];
    int G;
};

// sums[g][c] += {sum v, sum v*w} over the rows of segment g in this CTA's chunk
// mode 0 (forward statistics): v = x, w = x.   mode 1 (backward): v = d = dz * lrelu'(pre), w = xhat
__global__ void __launch_bounds__(SB_THREADS)
k_bn1d_seg_reduce(const float* __restrict__ x, const float* __restrict__ dz, int C, SegTable T, int chunk, int mode,
                  const float* __restrict__ mean_rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                  float slope, double* __restrict__ sums)
{
    __shared__ float4 red[2][SB_THREADS];
    const int g = blockIdx.y;
    const int r0 = T.off[g] + blockIdx.x * chunk;
    const int r1 = min(T.off[g + 1], r0 + chunk);
    if (r0 >= r1) return;
    const int c4n = C >> 2, lanes = SB_THREADS / c4n;
    const int c4 = threadIdx.x % c4n, lane = threadIdx.x / c4n;
    float m[4] = {0, 0, 0, 0}, rs[4] = {1, 1, 1, 1}, ga[4] = {1, 1, 1, 1}, be[4] = {0, 0, 0, 0};
    if (mode == 1) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = c4 * 4 + k;
            m[k] = mean_rstd[((size_t)g * C + c) * 2];
            rs[k] = mean_rstd[((size_t)g * C + c) * 2 + 1];
            ga[k] = __ldg(gamma + c);
            be[k] = __ldg(beta + c);
        }
    }
    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
    if (lane < lanes) {
        for (int r = r0 + lane; r < r1; r += lanes) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(x + (size_t)r * C) + c4);
            if (mode == 0) {
                s1.x += v.x; s1.y += v.y; s1.z += v.z; s1.w += v.w;
                s2.x = fmaf(v.x, v.x, s2.x); s2.y = fmaf(v.y, v.y, s2.y); s2.z = fmaf(v.z, v.z, s2.z); s2.w = fmaf(v.w, v.w, s2.w);
            } else {
                float4 d = __ldg(reinterpret_cast<const float4*>(dz + (size_t)r * C) + c4);
                const float h0 = (v.x - m[0]) * rs[0], h1 = (v.y - m[1]) * rs[1], h2 = (v.z - m[2]) * rs[2], h3 = (v.w - m[3]) * rs[3];
                if (slope >= 0.f) {
                    d.x *= fmaf(ga[0], h0, be[0]) > 0.f ? 1.f : slope;
                    d.y *= fmaf(ga[1], h1, be[1]) > 0.f ? 1.f : slope;
                    d.z *= fmaf(ga[2], h2, be[2]) > 0.f ? 1.f : slope;
                    d.w *= fmaf(ga[3], h3, be[3]) > 0.f ? 1.f : slope;
                }
                s1.x += d.x; s1.y += d.y; s1.z += d.z; s1.w += d.w;
                s2.x = fmaf(d.x, h0, s2.x); s2.y = fmaf(d.y, h1, s2.y); s2.z = fmaf(d.z, h2, s2.z); s2.w = fmaf(d.w, h3, s2.w);
            }
        }
    }
    red[0][threadIdx.x] = s1;
    red[1][threadIdx.x] = s2;
    __syncthreads();
    if (lane == 0) {
        for (int l = 1; l < lanes; ++l) {
            const float4 a = red[0][l * c4n + c4], q = red[1][l * c4n + c4];
            s1.x += a.x; s1.y += a.y; s1.z += a.z; s1.w += a.w;
            s2.x += q.x; s2.y += q.y; s2.z += q.z; s2.w += q.w;
        }
        double* dst = sums + ((size_t)g * C + c4 * 4) * 2;
        atomicAdd(dst + 0, (double)s1.x); atomicAdd(dst + 1, (double)s2.x);
        atomicAdd(dst + 2, (double)s1.y); atomicAdd(dst + 3, (double)s2.y);
        atomicAdd(dst + 4, (double)s1.z); atomicAdd(dst + 5, (double)s2.z);
        atomicAdd(dst + 6, (double)s1.w); atomicAdd(dst + 7, (double)s2.w);
    }
}

// z = lrelu(gamma * (x - mean) * rstd + beta); smem 3*C floats
__global__ void __launch_bounds__(SB_THREADS)
k_bn1d_seg_apply(const float* __restrict__ x, int C, SegTable T, int chunk, const double* __restrict__ stats,
                 const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ running_mean,
                 float* __restrict__ running_var, long long* __restrict__ nbt, float eps, float momentum, float slope,
                 float* __restrict__ z, float* __restrict__ mean_rstd)
{
    extern __shared__ float sm[];
    float* s_mean = sm;
    float* s_rs = sm + C;
    float* s_beta = sm + 2 * C;
    const int g = blockIdx.y;
    const double n = (double)(T.off[g + 1] - T.off[g]);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float mean, rstd;
        if (stats != nullptr) {
            const double mu = stats[((size_t)g * C + c) * 2] / n;
            double var = stats[((size_t)g * C + c) * 2 + 1] / n - mu * mu;
            if (var < 0) var = 0;
            mean = (float)mu;
            rstd = (float)(1.0 / sqrt(var + (double)eps));
        } else {
            mean = running_mean[c];
            rstd = 1.0f / sqrtf(running_var[c] + eps);
        }
        s_mean[c] = mean;
        s_rs[c] = rstd * __ldg(gamma + c);
        s_beta[c] = __ldg(beta + c);
        if (mean_rstd != nullptr && blockIdx.x == 0) {
            mean_rstd[((size_t)g * C + c) * 2] = mean;
            mean_rstd[((size_t)g * C + c) * 2 + 1] = rstd;
        }
    }
    // running statistics move once per frame, frame after frame (the reference's per-frame encoder calls)
    if (stats != nullptr && running_mean != nullptr && blockIdx.x == 0 && g == 0) {
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            float rm = running_mean[c], rv = running_var[c];
            for (int gg = 0; gg < T.G; ++gg) {
                const double ng = (double)(T.off[gg + 1] - T.off[gg]);
                if (ng <= 0) continue;
                const double mu = stats[((size_t)gg * C + c) * 2] / ng;
                double var = stats[((size_t)gg * C + c) * 2 + 1] / ng - mu * mu;
                if (var < 0) var = 0;
                const float vu = (float)(ng > 1 ? var * (ng / (ng - 1.0)) : var);
                rm = (1.f - momentum) * rm + momentum * (float)mu;
                rv = (1.f - momentum) * rv + momentum * vu;
            }
            running_mean[c] = rm;
            running_var[c] = rv;
        }
        if (threadIdx.x == 0 && nbt != nullptr) {
            int live = 0;
            for (int gg = 0; gg < T.G; ++gg) live += T.off[gg + 1] > T.off[gg];
            *nbt += live;
        }
    }
    __syncthreads();
    const int c4n = C >> 2;
    const int r0 = T.off[g] + blockIdx.x * chunk;
    const int r1 = min(T.off[g + 1], r0 + chunk);
    if (r0 >= r1) return;
    const int items = (r1 - r0) * c4n;
    const float4* x4 = reinterpret_cast<const float4*>(x) + (size_t)r0 * c4n;
    float4* z4 = reinterpret_cast<float4*>(z) + (size_t)r0 * c4n;
    for (int i = threadIdx.x; i < items; i += blockDim.x) {
        const int c = (i % c4n) * 4;
        const float4 v = __ldg(x4 + i);
        float4 o;
        o.x = (v.x - s_mean[c]) * s_rs[c] + s_beta[c];
        o.y = (v.y - s_mean[c + 1]) * s_rs[c + 1] + s_beta[c + 1];
        o.z = (v.z - s_mean[c + 2]) * s_rs[c + 2] + s_beta[c + 2];
        o.w = (v.w - s_mean[c + 3]) * s_rs[c + 3] + s_beta[c + 3];
        if (slope >= 0.f) {
            o.x = o.x > 0.f ? o.x : o.x * slope; o.y = o.y > 0.f ? o.y : o.y * slope;
            o.z = o.z > 0.f ? o.z : o.z * slope; o.w = o.w > 0.f ? o.w : o.w * slope;
        }
        z4[i] = o;
    }
}

// dx = gamma*rstd * (d - mean(d) - xhat*mean(d*xhat)) (batch statistics) or gamma*rstd*d (running statistics);
// dgamma / dbeta over all segments.  smem 6*C floats
__global__ void __launch_bounds__(SB_THREADS)
k_bn1d_seg_bwd_apply(const float* __restrict__ dz, const float* __restrict__ x, int C, SegTable T, int chunk,
                     const float* __restrict__ mean_rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                     const double* __restrict__ sums, float slope, int batch_stats, float* __restrict__ dx,
                     float* __restrict__ dgamma, float* __restrict__ dbeta)
{
    extern __shared__ float sm[];
    float* s_mean = sm;
    float* s_rstd = sm + C;
    float* s_g = sm + 2 * C;
    float* s_b = sm + 3 * C;
    float* s_m1 = sm + 4 * C;
    float* s_m2 = sm + 5 * C;
    const int g = blockIdx.y;
    const double n = (double)(T.off[g + 1] - T.off[g]);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        s_mean[c] = mean_rstd[((size_t)g * C + c) * 2];
        s_rstd[c] = mean_rstd[((size_t)g * C + c) * 2 + 1];
        s_g[c] = __ldg(gamma + c);
        s_b[c] = __ldg(beta + c);
        s_m1[c] = (batch_stats && n > 0) ? (float)(sums[((size_t)g * C + c) * 2] / n) : 0.f;
        s_m2[c] = (batch_stats && n > 0) ? (float)(sums[((size_t)g * C + c) * 2 + 1] / n) : 0.f;
    }
    if (blockIdx.x == 0 && g == 0) {
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            double sb = 0, sg = 0;
            for (int gg = 0; gg < T.G; ++gg) {
                sb += sums[((size_t)gg * C + c) * 2];
                sg += sums[((size_t)gg * C + c) * 2 + 1];
            }
            if (dbeta) dbeta[c] = (float)sb;
            if (dgamma) dgamma[c] = (float)sg;
        }
    }
    __syncthreads();
    const int c4n = C >> 2;
    const int r0 = T.off[g] + blockIdx.x * chunk;
    const int r1 = min(T.off[g + 1], r0 + chunk);
    if (r0 >= r1) return;
    const int items = (r1 - r0) * c4n;
    const float4* d4 = reinterpret_cast<const float4*>(dz) + (size_t)r0 * c4n;
    const float4* x4 = reinterpret_cast<const float4*>(x) + (size_t)r0 * c4n;
    float4* o4 = reinterpret_cast<float4*>(dx) + (size_t)r0 * c4n;
    for (int i = threadIdx.x; i < items; i += blockDim.x) {
        const int c = (i % c4n) * 4;
        const float4 v = __ldg(x4 + i);
        float4 d = __ldg(d4 + i);
        float h[4] = {(v.x - s_mean[c]) * s_rstd[c], (v.y - s_mean[c + 1]) * s_rstd[c + 1], (v.z - s_mean[c + 2]) * s_rstd[c + 2],
                      (v.w - s_mean[c + 3]) * s_rstd[c + 3]};
        float dd[4] = {d.x, d.y, d.z, d.w};
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (slope >= 0.f) dd[k] *= fmaf(s_g[c + k], h[k], s_b[c + k]) > 0.f ? 1.f : slope;
            o[k] = s_g[c + k] * s_rstd[c + k] * (dd[k] - s_m1[c + k] - h[k] * s_m2[c + k]);
        }
        o4[i] = make_float4(o[0], o[1], o[2], o[3]);
    }
}

static int make_table(const int* seg_rows_host, int G, SegTable* T, int* max_rows)
{
    if (G < 1 || G > SB_MAX_SEG) return 1;
    T->G = G;
    T->off[0] = 0;
    *max_rows = 0;
    for (int g = 0; g < G; ++g) {
        if (seg_rows_host[g] < 0) return 1;
        T->off[g + 1] = T->off[g] + seg_rows_host[g];
        if (seg_rows_host[g] > *max_rows) *max_rows = seg_rows_host[g];
    }
    for (int g = G + 1; g <= SB_MAX_SEG; ++g) T->off[g] = T->off[G];
    return 0;
}

// rows per CTA: enough CTAs to fill the machine, at most 32 rows per thread of the reduction
static int pick_rows(int max_rows, int c4n, int G)
{
    const int lanes = SB_THREADS / c4n;
    int chunk = lanes * 32;
    while (chunk > lanes * 4 && (long long)cdiv(max_rows, chunk) * G < 2 * 148) chunk >>= 1;
    return chunk;
}

static int bad(const char* what)
{
    set_last_error(what, cudaErrorInvalidValue);
    return (int)cudaErrorInvalidValue;
}

}  // namespace
}  // namespace rslo

using namespace rslo;

extern "C" int rslo_bn1d_seg_forward(const float* x, int C, const int* seg_rows_host, int G, const float* gamma,
                                     const float* beta, float* running_mean, float* running_var,
                                     long long* num_batches_tracked, float eps, float momentum, int training, float slope,
                                     double* stats, float* z, float* mean_rstd, rslo_stream_t stream)
{
    SegTable T;
    int max_rows;
    const int c4n = C / 4;
    if (C % 4 != 0 || c4n < 1 || c4n > SB_THREADS || SB_THREADS % c4n != 0 || make_table(seg_rows_host, G, &T, &max_rows))
        return bad("rslo_bn1d_seg_forward: bad shape");
    if (training && stats == nullptr) return bad("rslo_bn1d_seg_forward: training needs the statistics scratch");
    if (!training && running_mean == nullptr) return bad("rslo_bn1d_seg_forward: no statistics");
    if (max_rows == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int chunk = pick_rows(max_rows, c4n, G);
    const dim3 grid(cdiv(max_rows, chunk), G);
    if (training) {
        RSLO_COUNT();
        k_bn1d_seg_reduce<<<grid, SB_THREADS, 0, st>>>(x, nullptr, C, T, chunk, 0, nullptr, nullptr, nullptr, -1.f, stats);
        RSLO_CHECK_LAUNCH("rslo_bn1d_seg_forward(stats)");
    }
    RSLO_COUNT();
    k_bn1d_seg_apply<<<grid, SB_THREADS, 3 * C * sizeof(float), st>>>(x, C, T, chunk, training ? stats : nullptr, gamma, beta,
                                                                     running_mean, running_var, num_batches_tracked, eps,
                                                                     momentum, slope, z, mean_rstd);
    RSLO_CHECK_LAUNCH("rslo_bn1d_seg_forward(apply)");
    return 0;
}

extern "C" int rslo_bn1d_seg_backward(const float* dz, const float* x, int C, const int* seg_rows_host, int G,
                                      const float* mean_rstd, const float* gamma, const float* beta, float slope,
                                      int batch_stats, double* sums, float* dx, float* dgamma, float* dbeta,
                                      rslo_stream_t stream)
{
    SegTable T;
    int max_rows;
    const int c4n = C / 4;
    if (C % 4 != 0 || c4n < 1 || c4n > SB_THREADS || SB_THREADS % c4n != 0 || make_table(seg_rows_host, G, &T, &max_rows))
        return bad("rslo_bn1d_seg_backward: bad shape");
    if (max_rows == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int chunk = pick_rows(max_rows, c4n, G);
    const dim3 grid(cdiv(max_rows, chunk), G);
    RSLO_COUNT();
    k_bn1d_seg_reduce<<<grid, SB_THREADS, 0, st>>>(x, dz, C, T, chunk, 1, mean_rstd, gamma, beta, slope, sums);
    RSLO_CHECK_LAUNCH("rslo_bn1d_seg_backward(reduce)");
    RSLO_COUNT();
    k_bn1d_seg_bwd_apply<<<grid, SB_THREADS, 6 * C * sizeof(float), st>>>(dz, x, C, T, chunk, mean_rstd, gamma, beta, sums, slope,
                                                                         batch_stats, dx, dgamma, dbeta);
    RSLO_CHECK_LAUNCH("rslo_bn1d_seg_backward(apply)");
    return 0;
}
