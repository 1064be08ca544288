// K4: sparse 3-D convolution over output-stationary neighbour tables (sm_100a), FP32.
//
//   out[o,:] = act(scale * (bias + sum_k in[nbr[o,k],:] @ W[k]) + shift)
//
// Replaces spconv's per-offset gather -> cuBLAS GEMM -> scatter-add (27 GEMMs + 54 gather/scatter
// launches per layer) by ONE launch per layer with the bias / BatchNorm-affine / LeakyReLU epilogue
// fused, no atomics and a fixed summation order (k ascending, then input channel ascending), so the
// result is deterministic.  One warp owns one output row: the row's K table entries are fetched
// with one coalesced load, absent offsets are skipped warp-uniformly, each present neighbour row is
// read with one coalesced 128/256 B load and broadcast through shuffles; W[k] is read through L1
// (the whole filter bank is <= 442 KB and stays L1/L2 resident).
// FP32 FFMA on purpose: north_star asks for 1e-4 relative pose/loss parity, which rules out plain
// TF32/BF16 tensor-core inputs here.
#include "common.cuh"

namespace rslo {
namespace {

// Narrow outputs (COUT <= 16) would leave most lanes of a warp-per-row kernel idle in the FMA loop, so the warp is cut
// into GROUPS = 32 / GL lane groups (GL = 16 lanes for COUT <= 16, 8 for COUT <= 8) that take DIFFERENT neighbours of the
// row at the same time - group g the g-th, (g + GROUPS)-th, ... valid offset - and the group sums are added in a fixed
// order at the end (deterministic).  COUT >= 32: one group, the plain k-ascending order.
template <int COUT>
struct FwdGroups {
    static constexpr int GL = COUT > 16 ? 32 : (COUT > 8 ? 16 : 8);
    static constexpr int GROUPS = 32 / GL;
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(256)
k_spconv_fwd(const float* __restrict__ in, const int* __restrict__ nbr, int n_cap, const int* n_dev, int K,
             const float* __restrict__ W, const float* __restrict__ bias, const float* __restrict__ scale,
             const float* __restrict__ shift, int act, float slope, float* __restrict__ out)
{
    constexpr int GL = FwdGroups<COUT>::GL, GROUPS = FwdGroups<COUT>::GROUPS;
    constexpr int NA = (CIN + GL - 1) / GL, NO = (COUT + GL - 1) / GL;
    const int lane = threadIdx.x & 31;
    const int gl = lane & (GL - 1), grp = lane / GL;
    const int row = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (row >= dev_count(n_dev, n_cap)) return;
    float acc[NO];
#pragma unroll
    for (int t = 0; t < NO; ++t) acc[t] = 0.f;
    const int mynb = lane < K ? __ldg(nbr + (size_t)row * K + lane) : -1;
    unsigned valid = __ballot_sync(0xffffffffu, mynb >= 0);
    while (valid) {
        // this group's neighbour: the grp-th set bit of `valid` (none left: idles through the iteration with a = 0)
        unsigned mine = valid;
#pragma unroll
        for (int i = 0; i < GROUPS - 1; ++i)
            if (i < grp) mine &= mine - 1;
        const bool has = mine != 0;
        const int k = has ? __ffs(mine) - 1 : 0;
#pragma unroll
        for (int i = 0; i < GROUPS; ++i) valid &= valid - 1;               // (x & (x - 1) of 0 stays 0)
        const int j = __shfl_sync(0xffffffffu, mynb, k);
        const float* src = in + (size_t)(has ? j : 0) * CIN;
        float a[NA];
#pragma unroll
        for (int t = 0; t < NA; ++t) a[t] = (has && gl + GL * t < CIN) ? __ldg(src + gl + GL * t) : 0.f;
        const float* w = W + (size_t)k * CIN * COUT;
#pragma unroll
        for (int c = 0; c < CIN; ++c) {
            const float av = __shfl_sync(0xffffffffu, a[c / GL], (c % GL) + GL * grp);
#pragma unroll
            for (int t = 0; t < NO; ++t)
                if (COUT % GL == 0 || gl + GL * t < COUT)
                    acc[t] = fmaf(av, __ldg(w + c * COUT + gl + GL * t), acc[t]);
        }
    }
#pragma unroll
    for (int off = GL; off < 32; off <<= 1)
#pragma unroll
        for (int t = 0; t < NO; ++t) acc[t] += __shfl_xor_sync(0xffffffffu, acc[t], off);
    if (grp == 0) {
#pragma unroll
        for (int t = 0; t < NO; ++t) {
            const int co = gl + GL * t;
            if (co < COUT) {
                float v = acc[t];
                if (bias) v += __ldg(bias + co);
                if (scale) v = fmaf(v, __ldg(scale + co), __ldg(shift + co));
                if (act == 1) v = v > 0.f ? v : v * slope;
                out[(size_t)row * COUT + co] = v;
            }
        }
    }
}

// W [K,Cin,Cout] -> Wt [K,Cout,Cin] with optional offset mirroring (k -> K-1-k).
__global__ void k_transpose_w(const float* __restrict__ W, int K, int Cin, int Cout, int mirror,
                              float* __restrict__ Wt)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int tot = K * Cin * Cout;
    if (t >= tot) return;
    int ci = t % Cin, co = (t / Cin) % Cout, k = t / (Cin * Cout);
    int ks = mirror ? K - 1 - k : k;
    Wt[t] = W[((size_t)ks * Cin + ci) * Cout + co];
}

// dW[k] += sum over rows o of a chunk with nbr[o,k] >= 0 of in[nbr[o,k],:]^T (x) g[o,:]
// grid (chunks, K); 256 threads; each thread owns Cin*Cout/256 (>=1) accumulators.
constexpr int WG_ROWS = 1024;  // rows per chunk
constexpr int WG_TILE = 64;    // valid rows staged per step

template <int CIN, int COUT>
__global__ void __launch_bounds__(256)
k_spconv_wgrad(const float* __restrict__ in, const float* __restrict__ g, const int* __restrict__ nbr,
               int n_cap, const int* n_dev, int K, float* __restrict__ dW)
{
    // register tile per thread: TM input channels x TN output channels on a fixed 16 x 16 thread grid
    constexpr int TM = CIN >= 64 ? 4 : (CIN >= 32 ? 2 : 1);
    constexpr int TN = COUT >= 64 ? 4 : (COUT >= 32 ? 2 : 1);
    __shared__ int s_list_o[WG_ROWS];
    __shared__ int s_list_j[WG_ROWS];
    __shared__ int s_count;
    __shared__ __align__(16) float s_a[WG_TILE][CIN];
    __shared__ __align__(16) float s_g[WG_TILE][COUT];
    const int n = dev_count(n_dev, n_cap);
    const int k = blockIdx.y;
    const int row0 = blockIdx.x * WG_ROWS;
    if (row0 >= n) return;
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    // compact the chunk's valid (o, j) pairs
    for (int base = 0; base < WG_ROWS; base += 256) {
        int o = row0 + base + threadIdx.x;
        int j = (o < n) ? __ldg(nbr + (size_t)o * K + k) : -1;
        unsigned m = __ballot_sync(0xffffffffu, j >= 0);
        int lane = threadIdx.x & 31;
        int pos = 0;
        if (lane == 0 && m) pos = atomicAdd(&s_count, __popc(m));
        pos = __shfl_sync(0xffffffffu, pos, 0);
        if (j >= 0) {
            int p = pos + __popc(m & ((1u << lane) - 1));
            s_list_o[p] = o;
            s_list_j[p] = j;
        }
    }
    __syncthreads();
    const int cnt = s_count;
    if (cnt == 0) return;
    const int ci0 = (threadIdx.x >> 4) * TM, co0 = (threadIdx.x & 15) * TN;
    const bool active = ci0 < CIN && co0 < COUT;
    float acc[TM][TN];
#pragma unroll
    for (int a = 0; a < TM; ++a)
#pragma unroll
        for (int b = 0; b < TN; ++b) acc[a][b] = 0.f;
    for (int base = 0; base < cnt; base += WG_TILE) {
        const int m = min(WG_TILE, cnt - base);
        for (int t = threadIdx.x; t < m * CIN; t += 256) {
            int r = t / CIN, c = t % CIN;
            s_a[r][c] = __ldg(in + (size_t)s_list_j[base + r] * CIN + c);
        }
        for (int t = threadIdx.x; t < m * COUT; t += 256) {
            int r = t / COUT, c = t % COUT;
            s_g[r][c] = __ldg(g + (size_t)s_list_o[base + r] * COUT + c);
        }
        __syncthreads();
        if (active) {
#pragma unroll 4
            for (int r = 0; r < m; ++r) {
                float av[TM], gv[TN];
#pragma unroll
                for (int a = 0; a < TM; ++a) av[a] = s_a[r][ci0 + a];
#pragma unroll
                for (int b = 0; b < TN; ++b) gv[b] = s_g[r][co0 + b];
#pragma unroll
                for (int a = 0; a < TM; ++a)
#pragma unroll
                    for (int b = 0; b < TN; ++b) acc[a][b] = fmaf(av[a], gv[b], acc[a][b]);
            }
        }
        __syncthreads();
    }
    if (active) {
#pragma unroll
        for (int a = 0; a < TM; ++a)
#pragma unroll
            for (int b = 0; b < TN; ++b)
                atomicAdd(dW + ((size_t)k * CIN + ci0 + a) * COUT + co0 + b, acc[a][b]);
    }
}

// g_act[o,c] = g[o,c] * (out[o,c] > 0 ? 1 : slope)  (LeakyReLU backward from the saved OUTPUT; valid
// because slope > 0 keeps the sign) fused with the bias gradient dbias[c] += sum_o g_act[o,c].
// Thread (c, rl) walks rows rl, rl + rstep, ... so every global access is a coalesced row segment.
__global__ void __launch_bounds__(256)
k_act_bwd_colsum(const float* __restrict__ g, const float* __restrict__ out, int n_cap, const int* n_dev, int C,
                 int act, float slope, float* __restrict__ g_act, float* __restrict__ dbias)
{
    const int n = dev_count(n_dev, n_cap);
    const int rstep = 256 / C;
    const int c = threadIdx.x % C, rl = threadIdx.x / C;
    float s = 0.f;
    if (rl < rstep) {
        for (int o = blockIdx.x * rstep + rl; o < n; o += gridDim.x * rstep) {
            float v = g[(size_t)o * C + c];
            if (act == 1) {
                v = out[(size_t)o * C + c] > 0.f ? v : v * slope;
                g_act[(size_t)o * C + c] = v;
            }
            s += v;
        }
    }
    if (dbias) {
        __shared__ float red[256];
        red[threadIdx.x] = rl < rstep ? s : 0.f;
        __syncthreads();
        if (threadIdx.x < C) {
            float t = 0.f;
            for (int r = 0; r < rstep; ++r) t += red[r * C + threadIdx.x];
            atomicAdd(dbias + threadIdx.x, t);
        }
    }
}

__global__ void k_dense(const float* __restrict__ feat, int C, const uint2* __restrict__ cells,
                        const int* __restrict__ perm, int ncell, float* __restrict__ dense)
{
    int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= ncell) return;
    int row = site_lookup(cells, perm, (unsigned)cell);
    // dense [C, D*H*W]: consecutive threads write consecutive cells of one channel plane
    if (row < 0) {
        for (int c = 0; c < C; ++c) dense[(size_t)c * ncell + cell] = 0.f;
    } else {
        const float* f = feat + (size_t)row * C;
        for (int c = 0; c < C; ++c) dense[(size_t)c * ncell + cell] = __ldg(f + c);
    }
}

__global__ void k_dense_bwd(const float* __restrict__ gd, int C, const int* __restrict__ coors, int stride,
                            int n_cap, const int* n_dev, int H, int W, int ncell, float* __restrict__ gf)
{
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int r = (int)(t / C), c = (int)(t % C);
    if (r >= dev_count(n_dev, n_cap)) return;
    const int* co = coors + (size_t)r * stride + (stride - 3);
    int cell = (co[0] * H + co[1]) * W + co[2];
    gf[t] = gd[(size_t)c * ncell + cell];
}

#define RSLO_DISPATCH_CC(CIN_, COUT_, ...)                                     \
    if (Cin == CIN_ && Cout == COUT_) {                                        \
        constexpr int CI = CIN_, CO = COUT_;                                   \
        __VA_ARGS__;                                                           \
        handled = true;                                                        \
    }

#define RSLO_ALL_SHAPES(...)                                                   \
    RSLO_DISPATCH_CC(7, 16, __VA_ARGS__)                                       \
    RSLO_DISPATCH_CC(16, 7, __VA_ARGS__)                                       \
    RSLO_DISPATCH_CC(16, 16, __VA_ARGS__)                                      \
    RSLO_DISPATCH_CC(16, 32, __VA_ARGS__)                                      \
    RSLO_DISPATCH_CC(32, 16, __VA_ARGS__)                                      \
    RSLO_DISPATCH_CC(32, 32, __VA_ARGS__)                                      \
    RSLO_DISPATCH_CC(32, 64, __VA_ARGS__)                                      \
    RSLO_DISPATCH_CC(64, 32, __VA_ARGS__)                                      \
    RSLO_DISPATCH_CC(64, 64, __VA_ARGS__)

int launch_fwd(const float* in, const int* nbr, int n_cap, const int* n_dev, int K, int Cin, int Cout,
               const float* W, const float* bias, const float* scale, const float* shift, int act,
               float slope, float* out, cudaStream_t st)
{
    if (n_cap <= 0) return 0;
    if (K > 32 || K < 1) {
        set_last_error("spconv: K must be in 1..32", cudaErrorInvalidValue);
        return (int)cudaErrorInvalidValue;
    }
    const int grid = cdiv((long long)n_cap * 32, 256);
    bool handled = false;
    RSLO_COUNT();
    RSLO_ALL_SHAPES((k_spconv_fwd<CI, CO><<<grid, 256, 0, st>>>(in, nbr, n_cap, n_dev, K, W, bias, scale,
                                                               shift, act, slope, out)))
    if (!handled) {
        set_last_error("spconv: unsupported (Cin,Cout)", cudaErrorInvalidValue);
        return (int)cudaErrorInvalidValue;
    }
    RSLO_CHECK_LAUNCH("rslo_spconv");
    return 0;
}

}  // namespace
}  // namespace rslo

using namespace rslo;

extern "C" int rslo_spconv_forward(const float* in, const int32_t* nbr, int n_out_cap,
                                   const int32_t* n_out_dev, int K, int Cin, int Cout, const float* weight,
                                   const float* bias, const float* scale, const float* shift, int act,
                                   float slope, float* out, rslo_stream_t stream)
{
    return launch_fwd(in, nbr, n_out_cap, n_out_dev, K, Cin, Cout, weight, bias, scale, shift, act, slope,
                      out, (cudaStream_t)stream);
}

extern "C" int rslo_spconv_transpose_weight(const float* weight, int K, int Cin, int Cout, int mirror,
                                            float* weight_t, rslo_stream_t stream)
{
    int tot = K * Cin * Cout;
    RSLO_COUNT();
    k_transpose_w<<<cdiv(tot, 256), 256, 0, (cudaStream_t)stream>>>(weight, K, Cin, Cout, mirror, weight_t);
    RSLO_CHECK_LAUNCH("rslo_spconv_transpose_weight");
    return 0;
}

extern "C" int rslo_spconv_backward_data(const float* grad_out, const int32_t* nbr_t, int n_in_cap,
                                         const int32_t* n_in_dev, int K, int Cin, int Cout,
                                         const float* weight_t, float* grad_in, rslo_stream_t stream)
{
    // the data gradient IS a gather-convolution over the transposed table with W[k]^T
    return launch_fwd(grad_out, nbr_t, n_in_cap, n_in_dev, K, Cout, Cin, weight_t, nullptr, nullptr, nullptr,
                      0, 0.f, grad_in, (cudaStream_t)stream);
}

extern "C" int rslo_spconv_backward_weight(const float* in, const float* grad_out, const int32_t* nbr,
                                           int n_out_cap, const int32_t* n_out_dev, int K, int Cin,
                                           int Cout, float* grad_weight, float* grad_bias,
                                           rslo_stream_t stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    RSLO_CHECK(cudaMemsetAsync(grad_weight, 0, (size_t)K * Cin * Cout * sizeof(float), st));
    if (grad_bias) RSLO_CHECK(cudaMemsetAsync(grad_bias, 0, (size_t)Cout * sizeof(float), st));
    if (n_out_cap <= 0) return 0;
    dim3 grid(cdiv(n_out_cap, WG_ROWS), K);
    bool handled = false;
    RSLO_COUNT();
    RSLO_ALL_SHAPES((k_spconv_wgrad<CI, CO><<<grid, 256, 0, st>>>(in, grad_out, nbr, n_out_cap, n_out_dev, K,
                                                                 grad_weight)))
    if (!handled) {
        set_last_error("spconv wgrad: unsupported (Cin,Cout)", cudaErrorInvalidValue);
        return (int)cudaErrorInvalidValue;
    }
    if (grad_bias) {
        int blocks = cdiv(n_out_cap, 256 / Cout * 8);
        if (blocks > 148 * 4) blocks = 148 * 4;
        RSLO_COUNT();
        k_act_bwd_colsum<<<blocks, 256, 0, st>>>(grad_out, nullptr, n_out_cap, n_out_dev, Cout, 0, 0.f, nullptr,
                                                 grad_bias);
    }
    RSLO_CHECK_LAUNCH("rslo_spconv_backward_weight");
    return 0;
}

extern "C" int rslo_act_backward(const float* grad_out, const float* out, int n_cap, const int32_t* n_dev, int C,
                                 int act, float slope, float* grad_act, float* grad_bias, rslo_stream_t stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (C < 1 || C > 256 || (act == 1 && (!out || !grad_act))) {
        set_last_error("rslo_act_backward: bad argument", cudaErrorInvalidValue);
        return (int)cudaErrorInvalidValue;
    }
    if (grad_bias) RSLO_CHECK(cudaMemsetAsync(grad_bias, 0, (size_t)C * sizeof(float), st));
    if (n_cap <= 0) return 0;
    int blocks = cdiv(n_cap, 256 / C * 8);
    if (blocks > 148 * 4) blocks = 148 * 4;
    RSLO_COUNT();
    k_act_bwd_colsum<<<blocks, 256, 0, st>>>(grad_out, out, n_cap, n_dev, C, act, slope, grad_act, grad_bias);
    RSLO_CHECK_LAUNCH("rslo_act_backward");
    return 0;
}

extern "C" int rslo_dense_from_sites(const float* feat, int C, const uint32_t* cells, const int32_t* perm,
                                     int D, int H, int W, float* dense, rslo_stream_t stream)
{
    int ncell = D * H * W;
    RSLO_COUNT();
    k_dense<<<cdiv(ncell, 128), 128, 0, (cudaStream_t)stream>>>(feat, C, (const uint2*)cells, perm, ncell, dense);
    RSLO_CHECK_LAUNCH("rslo_dense_from_sites");
    return 0;
}

extern "C" int rslo_dense_backward(const float* grad_dense, int C, const int32_t* coors, int coor_stride,
                                   int n_cap, const int32_t* n_dev, int D, int H, int W, float* grad_feat,
                                   rslo_stream_t stream)
{
    if (n_cap <= 0) return 0;
    int ncell = D * H * W;
    RSLO_COUNT();
    k_dense_bwd<<<cdiv((long long)n_cap * C, 256), 256, 0, (cudaStream_t)stream>>>(
        grad_dense, C, coors, coor_stride, n_cap, n_dev, H, W, ncell, grad_feat);
    RSLO_CHECK_LAUNCH("rslo_dense_backward");
    return 0;
}
