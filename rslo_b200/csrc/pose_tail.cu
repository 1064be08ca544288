// The two "tails" of the dense head, each one kernel forward + one kernel backward (sm_100a).
//
// (1) k_head_tail_fwd / _bwd  —  `rslo/models/odom_pred.py:226-313` after the convolutions:
//       q normalisation (`:231-234`), masked spatial softmax confidences at temperature 1 and 20
//       (`rslo/layers/confidence.py:23-34`, `odom_pred.py:242-258`), per-cell local -> global (t,q)
//       (`rslo/data/dataset.py:121-208`, `rslo/utils/pose_utils.py:130-142`), confidence-weighted vote and pose
//       normalisation (`odom_pred.py:347-357,287-288`), pyramid masks: max-pool cascade of the input mask times the
//       average-pooled finer confidence (`odom_pred.py:210-216,262-264`), masked pyramid predictions (`:219-225`).
// (2) k_loss_tail_fwd / _bwd  —  `rslo/models/voxel_odom_net.py:727-795`: pseudo labels from the ICP residual
//       (`:727-735`, kornia rotation_matrix_to_quaternion), target (t,q) maps (`rslo/data/dataset.py:52-116`),
//       AdaptiveWeightedL2 on the pose and on the three pyramid levels (`rslo/core/losses.py:155-197`, focal_gamma 0).
//
// Forward kernels: one thread-block CLUSTER of 8 x 1024 threads per frame pair; every reduction (softmax max / sum,
// vote sums, loss sums) is a block reduction in double precision followed by an exchange of the CTA totals through
// distributed shared memory, combined in rank order -> deterministic; cluster barriers separate the passes (and
// order the global-memory maps one pass writes and the next pools).  Backward kernels: grid (chunks, B), the only
// reduction they need (softmax backward's sum) is cheap and read-only, so every CTA of a pair repeats it.
// Inputs from the trunk are NHWC with 32-channel rows (conv2d_tc.cu's narrow heads); outputs follow the reference's
// NCHW contract.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace rslo {
namespace {

constexpr int PT_THREADS = 1024;

struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 cross(const V3& a, const V3& b)
{
    return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ V3 operator+(const V3& a, const V3& b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(const V3& a, const V3& b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator*(float s, const V3& a) { return V3{s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ float dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// pose_utils.rotate_vec_by_q: t + 2 qs (qv x t) + 2 qv x (qv x t), in the reference's operation order
__device__ __forceinline__ V3 rotate_by_q(const V3& t, float qs, const V3& qv)
{
    const V3 b = cross(qv, t);
    const V3 c = 2.f * cross(qv, b);
    const V3 b2 = qs * (2.f * b);
    return (t + b2) + c;
}
// adjoint of rotate_by_q w.r.t. (t, qs, qv) for an output gradient g
__device__ __forceinline__ void rotate_by_q_bwd(const V3& t, float qs, const V3& qv, const V3& g, V3& dt, float& dqs, V3& dqv)
{
    const V3 b = cross(qv, t);
    const V3 gxq = cross(g, qv);
    const V3 db = 2.f * gxq;                                   // through c = 2 qv x b
    dt = (g + (2.f * qs) * gxq) + cross(db, qv);
    dqs = 2.f * dot(g, b);
    dqv = ((2.f * qs) * cross(t, g) + cross(t, db)) + 2.f * cross(b, g);
}

template <int N>
__device__ __forceinline__ void block_sum(double (&v)[N], double* sh /* [N][32] */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double x = v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) sh[i * 32 + warp] = x;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double x = lane < (int)(blockDim.x >> 5) ? sh[i * 32 + lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        v[i] = x;
    }
    __syncthreads();
}
__device__ __forceinline__ void block_max2(float& a, float& b, float* sh /* [2][32] */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, o));
        b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
    }
    if (lane == 0) {
        sh[warp] = a;
        sh[32 + warp] = b;
    }
    __syncthreads();
    a = lane < (int)(blockDim.x >> 5) ? sh[lane] : -INFINITY;
    b = lane < (int)(blockDim.x >> 5) ? sh[32 + lane] : -INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, o));
        b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
    }
    __syncthreads();
}

// ---- thread-block cluster variants: the PT_CLUSTER CTAs of one image each reduce their share, exchange the
// CTA totals through distributed shared memory and combine them in rank order (every CTA gets the same bits)
constexpr int PT_CLUSTER = 8;

template <int N>
__device__ __forceinline__ void cluster_sum(double (&v)[N], double* sh /* [N][32] */, double* xch /* [2][N] */)
{
    cg::cluster_group cl = cg::this_cluster();
    block_sum<N>(v, sh);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < N; ++i) xch[i] = v[i];
    }
    cl.sync();
    if (threadIdx.x < N) {
        double t = 0.0;
        for (unsigned r = 0; r < cl.num_blocks(); ++r) t += *cl.map_shared_rank(xch + threadIdx.x, r);
        xch[N + threadIdx.x] = t;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = xch[N + i];
    cl.sync();                                  // nobody overwrites xch while a neighbour still reads it
}
__device__ __forceinline__ void cluster_max2(float& a, float& b, float* sh /* [2][32] */, float* xch /* [4] */)
{
    cg::cluster_group cl = cg::this_cluster();
    block_max2(a, b, sh);
    if (threadIdx.x == 0) {
        xch[0] = a;
        xch[1] = b;
    }
    cl.sync();
    if (threadIdx.x < 2) {
        float t = -INFINITY;
        for (unsigned r = 0; r < cl.num_blocks(); ++r) t = fmaxf(t, *cl.map_shared_rank(xch + threadIdx.x, r));
        xch[2 + threadIdx.x] = t;
    }
    __syncthreads();
    a = xch[2];
    b = xch[3];
    cl.sync();
}

// cell anchor of pixel (i, j) of a map that is `s` times coarser than the geometry's grid
__device__ __forceinline__ V3 anchor(const rslo_tq_geom_t& G, int i, int j)
{
    return V3{((float)j - G.ox) * G.vsx, (-(float)i + G.oy) * G.vsy, (0.f - G.oz) * G.vsz};
}

struct PixelTQ {
    V3 tl, p, tg;
    float qn[4], qg[4];
    float nq, n2;
};
__device__ __forceinline__ PixelTQ pixel_tq(const float* __restrict__ row, const rslo_tq_geom_t& G, int i, int j)
{
    PixelTQ o;
    const float4 a = __ldg(reinterpret_cast<const float4*>(row));
    const float4 c = __ldg(reinterpret_cast<const float4*>(row) + 1);
    o.tl = V3{a.x, a.y, a.z};
    const float q0 = a.w, q1 = c.x, q2 = c.y, q3 = c.z;
    o.nq = sqrtf(((q0 * q0 + q1 * q1) + q2 * q2) + q3 * q3);
    o.qn[0] = q0 / o.nq; o.qn[1] = q1 / o.nq; o.qn[2] = q2 / o.nq; o.qn[3] = q3 / o.nq;
    o.p = anchor(G, i, j);
    o.tg = rotate_by_q(o.tl - o.p, o.qn[0], V3{o.qn[1], o.qn[2], o.qn[3]}) + o.p;
    o.n2 = fmaxf(sqrtf(((o.qn[0] * o.qn[0] + o.qn[1] * o.qn[1]) + o.qn[2] * o.qn[2]) + o.qn[3] * o.qn[3]), 1e-12f);
#pragma unroll
    for (int k = 0; k < 4; ++k) o.qg[k] = o.qn[k] / o.n2;
    return o;
}

// saved per pair: [0..1] max_t, max_r; [2..5] sum_t1, sum_t20, sum_r1, sum_r20; [6..8] A_t; [9] S_t; [10..13] A_q; [14] S_r
constexpr int HT_SAVE = 16;

__global__ void __cluster_dims__(PT_CLUSTER, 1, 1) __launch_bounds__(PT_THREADS)
k_head_tail_fwd(const float* __restrict__ tq32, const float* __restrict__ tl32, const float* __restrict__ rl32,
                const float* __restrict__ mask, const float* __restrict__ py0_32, const float* __restrict__ py1_32,
                rslo_tq_geom_t G, float* __restrict__ pose_t, float* __restrict__ pose_q, float* __restrict__ tq_g,
                float* __restrict__ t_conf, float* __restrict__ r_conf, float* __restrict__ pm2_pred,
                float* __restrict__ pm2_mask, float* __restrict__ pm1_pred, float* __restrict__ pm1_mask,
                float* __restrict__ pm0_pred, float* __restrict__ pm0_mask, float* __restrict__ base1,
                float* __restrict__ base0, float* __restrict__ save)
{
    // grid (PT_CLUSTER, B): one cluster per image; CTA `crank` owns pixels crank*PT_THREADS + tid, + PT_CLUSTER*PT_THREADS, ...
    __shared__ double sh[9 * 32];
    __shared__ double xch_d[2 * 9];
    __shared__ float shf[64];
    __shared__ float xch_f[4];
    cg::cluster_group cl = cg::this_cluster();
    const int b = blockIdx.y, tid = threadIdx.x;
    const int p_first = blockIdx.x * PT_THREADS + tid, p_step = PT_CLUSTER * PT_THREADS;
    const int H = G.H, W = G.W, HW = H * W;
    const float* m = mask + (size_t)b * HW;
    const float* tl = tl32 + (size_t)b * HW * 32;
    const float* rl = rl32 + (size_t)b * HW * 32;
    const float* tq = tq32 + (size_t)b * HW * 32;

    // pass 1: maxima of the masked logits
    float mt = -INFINITY, mr = -INFINITY;
    for (int p = p_first; p < HW; p += p_step) {
        const bool on = __ldg(m + p) > 0.f;
        mt = fmaxf(mt, on ? __ldg(tl + (size_t)p * 32) : -1000.f);
        mr = fmaxf(mr, on ? __ldg(rl + (size_t)p * 32) : -1000.f);
    }
    cluster_max2(mt, mr, shf, xch_f);
    const float mt20 = mt / 20.f, mr20 = mr / 20.f;
    // pass 2: softmax denominators at temperature 1 and 20
    double s[4] = {0, 0, 0, 0};
    for (int p = p_first; p < HW; p += p_step) {
        const bool on = __ldg(m + p) > 0.f;
        const float xt = on ? __ldg(tl + (size_t)p * 32) : -1000.f;
        const float xr = on ? __ldg(rl + (size_t)p * 32) : -1000.f;
        s[0] += (double)expf(xt - mt);
        s[1] += (double)expf(xt / 20.f - mt20);
        s[2] += (double)expf(xr - mr);
        s[3] += (double)expf(xr / 20.f - mr20);
    }
    cluster_sum<4>(s, sh, xch_d);
    const float st1 = (float)s[0], st20 = (float)s[1], sr1 = (float)s[2], sr20 = (float)s[3];
    // pass 3: confidences, global (t,q), vote sums, finest pyramid level
    double v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int p = p_first; p < HW; p += p_step) {
        const int i = p / W, j = p - i * W;
        const float mk = __ldg(m + p);
        const bool on = mk > 0.f;
        const float xt = on ? __ldg(tl + (size_t)p * 32) : -1000.f;
        const float xr = on ? __ldg(rl + (size_t)p * 32) : -1000.f;
        const float ct = expf(xt - mt) / st1, cr = expf(xr - mr) / sr1;
        const float ct20 = expf(xt / 20.f - mt20) / st20, cr20 = expf(xr / 20.f - mr20) / sr20;
        const PixelTQ q = pixel_tq(tq + (size_t)p * 32, G, i, j);
        t_conf[(size_t)b * HW + p] = ct;
        r_conf[(size_t)b * HW + p] = cr;
        float* g = tq_g + (size_t)b * 7 * HW + p;
        g[0] = q.tg.x * mk; g[HW] = q.tg.y * mk; g[2 * HW] = q.tg.z * mk;
        g[3 * HW] = q.qg[0] * mk; g[4 * HW] = q.qg[1] * mk; g[5 * HW] = q.qg[2] * mk; g[6 * HW] = q.qg[3] * mk;
        float* l = pm2_pred + (size_t)b * 7 * HW + p;
        l[0] = q.tl.x * mk; l[HW] = q.tl.y * mk; l[2 * HW] = q.tl.z * mk;
        l[3 * HW] = q.qn[0] * mk; l[4 * HW] = q.qn[1] * mk; l[5 * HW] = q.qn[2] * mk; l[6 * HW] = q.qn[3] * mk;
        pm2_mask[(size_t)b * 2 * HW + p] = mk * ct20;
        pm2_mask[(size_t)b * 2 * HW + HW + p] = mk * cr20;
        v[0] += (double)(q.tg.x * ct); v[1] += (double)(q.tg.y * ct); v[2] += (double)(q.tg.z * ct);
        v[3] += (double)ct;
        v[4] += (double)(q.qg[0] * cr); v[5] += (double)(q.qg[1] * cr); v[6] += (double)(q.qg[2] * cr);
        v[7] += (double)(q.qg[3] * cr);
        v[8] += (double)cr;
    }
    cluster_sum<9>(v, sh, xch_d);               // its cluster barriers also order the pm2_mask stores before the pooling
    if (tid == 0 && blockIdx.x == 0) {
        const float St = (float)v[3], Sr = (float)v[8];
        float* sv = save + (size_t)b * HT_SAVE;
        sv[0] = mt; sv[1] = mr; sv[2] = st1; sv[3] = st20; sv[4] = sr1; sv[5] = sr20;
        sv[6] = (float)v[0]; sv[7] = (float)v[1]; sv[8] = (float)v[2]; sv[9] = St;
        sv[10] = (float)v[4]; sv[11] = (float)v[5]; sv[12] = (float)v[6]; sv[13] = (float)v[7]; sv[14] = Sr;
        sv[15] = 0.f;
        pose_t[b * 3 + 0] = sv[6] / (St + 1e-12f);
        pose_t[b * 3 + 1] = sv[7] / (St + 1e-12f);
        pose_t[b * 3 + 2] = sv[8] / (St + 1e-12f);
        float qv[4];
        for (int k = 0; k < 4; ++k) qv[k] = sv[10 + k] / (Sr + 1e-12f);
        const float n = sqrtf(((qv[0] * qv[0] + qv[1] * qv[1]) + qv[2] * qv[2]) + qv[3] * qv[3]);
        for (int k = 0; k < 4; ++k) pose_q[b * 4 + k] = qv[k] / (n + 1e-12f);
    }
    // pyramid level 1 (H/2 x W/2) and level 0 (H/4 x W/4): MaxPool2d(3, 2, 1) of the occupancy mask, times
    // AvgPool2d(3, 2, 1) (zero padded, /9) of the finer level's confidence mask
    const float* fine_mask = pm2_mask + (size_t)b * 2 * HW;
    const float* fine_occ = m;
    int fh = H, fw = W;
    for (int lvl = 1; lvl >= 0; --lvl) {
        const int ch = fh / 2, cw = fw / 2, cn = ch * cw;
        float* occ = (lvl == 1 ? base1 : base0) + (size_t)b * cn;
        float* pmask = (lvl == 1 ? pm1_mask : pm0_mask) + (size_t)b * 2 * cn;
        float* ppred = (lvl == 1 ? pm1_pred : pm0_pred) + (size_t)b * 7 * cn;
        const float* praw = (lvl == 1 ? py1_32 : py0_32) + (size_t)b * cn * 32;
        for (int p = p_first; p < cn; p += p_step) {
            const int i = p / cw, j = p - i * cw;
            float mx = -INFINITY, a0 = 0.f, a1 = 0.f;
            for (int dy = -1; dy <= 1; ++dy) {
                const int y = 2 * i + dy;
                if (y < 0 || y >= fh) continue;
                for (int dx = -1; dx <= 1; ++dx) {
                    const int x = 2 * j + dx;
                    if (x < 0 || x >= fw) continue;
                    mx = fmaxf(mx, fine_occ[y * fw + x]);
                    a0 += fine_mask[y * fw + x];
                    a1 += fine_mask[fh * fw + y * fw + x];
                }
            }
            occ[p] = mx;
            pmask[p] = mx * (a0 / 9.f);
            pmask[cn + p] = mx * (a1 / 9.f);
            const float keep = mx > 0.f ? 1.f : 0.f;
            const float4 r0 = __ldg(reinterpret_cast<const float4*>(praw + (size_t)p * 32));
            const float4 r1 = __ldg(reinterpret_cast<const float4*>(praw + (size_t)p * 32) + 1);
            ppred[p] = r0.x * keep; ppred[cn + p] = r0.y * keep; ppred[2 * cn + p] = r0.z * keep; ppred[3 * cn + p] = r0.w * keep;
            ppred[4 * cn + p] = r1.x * keep; ppred[5 * cn + p] = r1.y * keep; ppred[6 * cn + p] = r1.z * keep;
        }
        cl.sync();                              // level-1 masks of the whole image before level 0 pools them
        fine_mask = pmask;
        fine_occ = occ;
        fh = ch;
        fw = cw;
    }
}

__global__ void __launch_bounds__(PT_THREADS)
k_head_tail_bwd(const float* __restrict__ tq32, const float* __restrict__ mask, const float* __restrict__ t_conf,
                const float* __restrict__ r_conf, const float* __restrict__ base1, const float* __restrict__ base0,
                const float* __restrict__ save, rslo_tq_geom_t G, const float* __restrict__ g_t,
                const float* __restrict__ g_q, const float* __restrict__ g_tqg, const float* __restrict__ g_tconf,
                const float* __restrict__ g_rconf, const float* __restrict__ g_pm2, const float* __restrict__ g_pm1,
                const float* __restrict__ g_pm0, float* __restrict__ d_tq32, float* __restrict__ d_tl32,
                float* __restrict__ d_rl32, float* __restrict__ d_py1_32, float* __restrict__ d_py0_32)
{
    __shared__ double sh[2 * 32];
    const int b = blockIdx.y, tid = threadIdx.x;
    const int H = G.H, W = G.W, HW = H * W;
    const float* m = mask + (size_t)b * HW;
    const float* tq = tq32 + (size_t)b * HW * 32;
    const float* sv = save + (size_t)b * HT_SAVE;
    // vote: t = A_t / (S_t + eps), rot = q_v / (|q_v| + eps), q_v = A_q / (S_r + eps)
    const float St = sv[9] + 1e-12f, Sr = sv[14] + 1e-12f;
    V3 dAt = V3{0.f, 0.f, 0.f};
    float dSt = 0.f, dAq[4] = {0.f, 0.f, 0.f, 0.f}, dSr = 0.f;
    if (g_t) {
        const V3 g = V3{g_t[b * 3], g_t[b * 3 + 1], g_t[b * 3 + 2]};
        dAt = (1.f / St) * g;
        dSt = -(sv[6] * g.x + sv[7] * g.y + sv[8] * g.z) / (St * St);
    }
    if (g_q) {
        float qv[4], g[4], dq[4];
        for (int k = 0; k < 4; ++k) {
            qv[k] = sv[10 + k] / Sr;
            g[k] = g_q[b * 4 + k];
        }
        const float n = sqrtf(((qv[0] * qv[0] + qv[1] * qv[1]) + qv[2] * qv[2]) + qv[3] * qv[3]);
        const float ne = n + 1e-12f;
        const float qg = qv[0] * g[0] + qv[1] * g[1] + qv[2] * g[2] + qv[3] * g[3];
        for (int k = 0; k < 4; ++k) dq[k] = g[k] / ne - qv[k] * qg / (fmaxf(n, 1e-30f) * ne * ne);
        float acc = 0.f;
        for (int k = 0; k < 4; ++k) {
            dAq[k] = dq[k] / Sr;
            acc += sv[10 + k] * dq[k];
        }
        dSr = -acc / (Sr * Sr);
    }
    // pass 1: softmax backward needs D = sum_p c(p) dc(p)
    double D[2] = {0, 0};
    for (int p = tid; p < HW; p += PT_THREADS) {
        const int i = p / W, j = p - i * W;
        const PixelTQ q = pixel_tq(tq + (size_t)p * 32, G, i, j);
        const float ct = __ldg(t_conf + (size_t)b * HW + p), cr = __ldg(r_conf + (size_t)b * HW + p);
        float dct = dot(q.tg, dAt) + dSt;
        float dcr = ((q.qg[0] * dAq[0] + q.qg[1] * dAq[1]) + q.qg[2] * dAq[2]) + q.qg[3] * dAq[3] + dSr;
        if (g_tconf) dct += __ldg(g_tconf + (size_t)b * HW + p);
        if (g_rconf) dcr += __ldg(g_rconf + (size_t)b * HW + p);
        D[0] += (double)(ct * dct);
        D[1] += (double)(cr * dcr);
    }
    block_sum<2>(D, sh);
    const float Dt = (float)D[0], Dr = (float)D[1];
    // pass 2: per-pixel gradients.  The grid is (chunks, B): every CTA of an image repeats the cheap read-only
    // pass 1 (same order -> the same D, bit for bit) and writes only its own share of the pixels, so the 6.5 MB of
    // 32-channel gradient rows per image are written by ~70 SMs instead of one.
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int pstride = gridDim.x * PT_THREADS;
    for (int p = blockIdx.x * PT_THREADS + tid; p < HW; p += pstride) {
        const int i = p / W, j = p - i * W;
        const float mk = __ldg(m + p);
        const bool on = mk > 0.f;
        const PixelTQ q = pixel_tq(tq + (size_t)p * 32, G, i, j);
        const float ct = __ldg(t_conf + (size_t)b * HW + p), cr = __ldg(r_conf + (size_t)b * HW + p);
        float dct = dot(q.tg, dAt) + dSt;
        float dcr = ((q.qg[0] * dAq[0] + q.qg[1] * dAq[1]) + q.qg[2] * dAq[2]) + q.qg[3] * dAq[3] + dSr;
        if (g_tconf) dct += __ldg(g_tconf + (size_t)b * HW + p);
        if (g_rconf) dcr += __ldg(g_rconf + (size_t)b * HW + p);
        const float dxt = on ? ct * (dct - Dt) : 0.f;
        const float dxr = on ? cr * (dcr - Dr) : 0.f;
        V3 dtg = ct * dAt;
        float dqg[4] = {cr * dAq[0], cr * dAq[1], cr * dAq[2], cr * dAq[3]};
        if (g_tqg) {
            const float* g = g_tqg + (size_t)b * 7 * HW + p;
            dtg = dtg + mk * V3{__ldg(g), __ldg(g + HW), __ldg(g + 2 * HW)};
#pragma unroll
            for (int k = 0; k < 4; ++k) dqg[k] += mk * __ldg(g + (3 + k) * HW);
        }
        // q_g = q_n / max(|q_n|, eps)
        const float qd = ((q.qg[0] * dqg[0] + q.qg[1] * dqg[1]) + q.qg[2] * dqg[2]) + q.qg[3] * dqg[3];
        float dqn[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) dqn[k] = (dqg[k] - q.qg[k] * qd) / q.n2;
        // t_g = rotate(t_l - p, q_n) + p
        V3 dtl, dqv;
        float dqs;
        rotate_by_q_bwd(q.tl - q.p, q.qn[0], V3{q.qn[1], q.qn[2], q.qn[3]}, dtg, dtl, dqs, dqv);
        dqn[0] += dqs; dqn[1] += dqv.x; dqn[2] += dqv.y; dqn[3] += dqv.z;
        if (g_pm2) {
            const float* g = g_pm2 + (size_t)b * 7 * HW + p;
            dtl = dtl + mk * V3{__ldg(g), __ldg(g + HW), __ldg(g + 2 * HW)};
#pragma unroll
            for (int k = 0; k < 4; ++k) dqn[k] += mk * __ldg(g + (3 + k) * HW);
        }
        // q_n = q_raw / |q_raw|
        const float qq = ((q.qn[0] * dqn[0] + q.qn[1] * dqn[1]) + q.qn[2] * dqn[2]) + q.qn[3] * dqn[3];
        float dqr[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) dqr[k] = (dqn[k] - q.qn[k] * qq) / q.nq;
        float4* o = reinterpret_cast<float4*>(d_tq32 + ((size_t)b * HW + p) * 32);
        o[0] = make_float4(dtl.x, dtl.y, dtl.z, dqr[0]);
        o[1] = make_float4(dqr[1], dqr[2], dqr[3], 0.f);
        float4* ot = reinterpret_cast<float4*>(d_tl32 + ((size_t)b * HW + p) * 32);
        float4* orr = reinterpret_cast<float4*>(d_rl32 + ((size_t)b * HW + p) * 32);
        ot[0] = make_float4(dxt, 0.f, 0.f, 0.f);
        orr[0] = make_float4(dxr, 0.f, 0.f, 0.f);
        ot[1] = zero4; orr[1] = zero4;
#pragma unroll
        for (int k = 2; k < 8; ++k) {
            o[k] = zero4;
            ot[k] = zero4;
            orr[k] = zero4;
        }
    }
    // pyramid predictions: d raw = g * [occupancy > 0]
    int cn = (H / 2) * (W / 2);
    for (int lvl = 1; lvl >= 0; --lvl) {
        const float* occ = (lvl == 1 ? base1 : base0) + (size_t)b * cn;
        const float* g = lvl == 1 ? g_pm1 : g_pm0;
        float* d = (lvl == 1 ? d_py1_32 : d_py0_32) + (size_t)b * cn * 32;
        for (int p = blockIdx.x * PT_THREADS + tid; p < cn; p += pstride) {
            const float keep = (g != nullptr && occ[p] > 0.f) ? 1.f : 0.f;
            float v[7];
#pragma unroll
            for (int k = 0; k < 7; ++k) v[k] = g ? keep * __ldg(g + (size_t)b * 7 * cn + (size_t)k * cn + p) : 0.f;
            float4* o = reinterpret_cast<float4*>(d + (size_t)p * 32);
            o[0] = make_float4(v[0], v[1], v[2], v[3]);
            o[1] = make_float4(v[4], v[5], v[6], 0.f);
#pragma unroll
            for (int k = 2; k < 8; ++k) o[k] = zero4;
        }
        cn /= 4;
    }
}

// ---------------------------------------------------------------------------------------------------
// loss tail
// ---------------------------------------------------------------------------------------------------
// kornia 0.4.0 quaternion_to_rotation_matrix on (x,y,z,w), L2-normalised with eps 1e-12; input here is (w,x,y,z)
__device__ void quat_wxyz_to_R(const float* q, float* R)
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
}
// kornia 0.4.0 rotation_matrix_to_quaternion (x,y,z,w), eps 1e-8, safe division by clamp(min = FLT_MIN); -> (w,x,y,z)
__device__ void R_to_quat_wxyz(const float* R, float* q)
{
    const float tiny = 1.17549435e-38f, eps = 1e-8f;
    const float m00 = R[0], m01 = R[1], m02 = R[2], m10 = R[3], m11 = R[4], m12 = R[5], m20 = R[6], m21 = R[7], m22 = R[8];
    const float trace = (m00 + m11) + m22;
    float x, y, z, w;
    if (trace > 0.f) {
        const float sq = sqrtf(trace + 1.f) * 2.f, d = fmaxf(sq, tiny);
        x = (m21 - m12) / d; y = (m02 - m20) / d; z = (m10 - m01) / d; w = 0.25f * sq;
    } else if (m00 > m11 && m00 > m22) {
        const float sq = sqrtf(((1.f + m00) - m11) - m22 + eps) * 2.f, d = fmaxf(sq, tiny);
        x = 0.25f * sq; y = (m01 + m10) / d; z = (m02 + m20) / d; w = (m21 - m12) / d;
    } else if (m11 > m22) {
        const float sq = sqrtf(((1.f + m11) - m00) - m22 + eps) * 2.f, d = fmaxf(sq, tiny);
        x = (m01 + m10) / d; y = 0.25f * sq; z = (m12 + m21) / d; w = (m02 - m20) / d;
    } else {
        const float sq = sqrtf(((1.f + m22) - m00) - m11 + eps) * 2.f, d = fmaxf(sq, tiny);
        x = (m02 + m20) / d; y = (m12 + m21) / d; z = 0.25f * sq; w = (m10 - m01) / d;
    }
    q[0] = w; q[1] = x; q[2] = y; q[3] = z;
}

// per pair: [0] L_T, [1] L_R, [2..4] LpyT (levels 0,1,2), [5..7] LpyR, [8..10] sum mask_t, [11..13] sum mask_r,
//           [14..16] t*, [17..20] q*
constexpr int LT_SAVE = 24;

struct PyrLevel {
    const float* pred;      // [B][7][h][w]
    const float* mask;      // [B][2][h][w]
    float* dpred;           // backward only
    int h, w, stride;
};

__device__ __forceinline__ void target_at(const rslo_tq_geom_t& G, int i, int j, const V3& ts, const float* qs_inv, V3& tl)
{
    const V3 p = anchor(G, i, j);
    tl = rotate_by_q(ts - p, qs_inv[0], V3{qs_inv[1], qs_inv[2], qs_inv[3]}) + p;
}

__global__ void __cluster_dims__(PT_CLUSTER, 1, 1) __launch_bounds__(PT_THREADS)
k_loss_tail_fwd(const float* __restrict__ T_pred, const float* __restrict__ q_pred, PyrLevel L0, PyrLevel L1, PyrLevel L2,
                const float* __restrict__ res_r, const float* __restrict__ res_t, int identity_pose, rslo_tq_geom_t G, int B,
                const float* __restrict__ alpha_t, const float* __restrict__ alpha_r, const float* __restrict__ alpha_pt,
                const float* __restrict__ alpha_pr, float w_t, float w_r, float w_pt, float w_pr, float* __restrict__ tq_target,
                float* __restrict__ save, float* __restrict__ losses /* [8] */, int* __restrict__ counter)
{
    // grid (PT_CLUSTER, B): one cluster per pair, the maps are shared out over its CTAs (see k_head_tail_fwd)
    __shared__ double sh[12 * 32];
    __shared__ double xch_d[2 * 12];
    __shared__ float s_lab[8];
    __shared__ int s_last;
    const int b = blockIdx.y, tid = threadIdx.x;
    const int p_first = blockIdx.x * PT_THREADS + tid, p_step = PT_CLUSTER * PT_THREADS;
    if (tid == 0) s_last = 0;
    const int H = G.H, W = G.W, HW = H * W;
    if (tid == 0) {
        float R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, T[3] = {0, 0, 0};
        if (!identity_pose) {
            quat_wxyz_to_R(q_pred + b * 4, R);
            for (int k = 0; k < 3; ++k) T[k] = T_pred[b * 3 + k];
        }
        const float* rr = res_r + b * 9;
        float Rs[9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) Rs[i * 3 + j] = (rr[i * 3] * R[j] + rr[i * 3 + 1] * R[3 + j]) + rr[i * 3 + 2] * R[6 + j];
        float q[4];
        R_to_quat_wxyz(Rs, q);
        const float sg = q[0] > 0.f ? 1.f : (q[0] < 0.f ? -1.f : 0.f);
        for (int k = 0; k < 4; ++k) s_lab[3 + k] = q[k] * sg;
        for (int i = 0; i < 3; ++i) s_lab[i] = ((rr[i * 3] * T[0] + rr[i * 3 + 1] * T[1]) + rr[i * 3 + 2] * T[2]) + res_t[b * 3 + i];
    }
    __syncthreads();
    const V3 ts = V3{s_lab[0], s_lab[1], s_lab[2]};
    const float qs[4] = {s_lab[3], s_lab[4], s_lab[5], s_lab[6]};
    const float qi[4] = {qs[0], -qs[1], -qs[2], -qs[3]};
    // target map at full resolution
    for (int p = p_first; p < HW; p += p_step) {
        const int i = p / W, j = p - i * W;
        V3 tl;
        target_at(G, i, j, ts, qi, tl);
        float* o = tq_target + (size_t)b * 7 * HW + p;
        o[0] = tl.x; o[HW] = tl.y; o[2 * HW] = tl.z;
        o[3 * HW] = qs[0]; o[4 * HW] = qs[1]; o[5 * HW] = qs[2]; o[6 * HW] = qs[3];
    }
    double acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};     // per level: sum_t, sum_r, mask_t, mask_r
    const PyrLevel Ls[3] = {L0, L1, L2};
#pragma unroll
    for (int l = 0; l < 3; ++l) {
        const PyrLevel& L = Ls[l];
        const int n = L.h * L.w;
        const float* pr = L.pred + (size_t)b * 7 * n;
        const float* mk = L.mask + (size_t)b * 2 * n;
        for (int p = p_first; p < n; p += p_step) {
            const int i = p / L.w, j = p - i * L.w;
            V3 tl;
            target_at(G, i * L.stride, j * L.stride, ts, qi, tl);
            const float mt = __ldg(mk + p), mr = __ldg(mk + n + p);
            const float d0 = __ldg(pr + p) - tl.x, d1 = __ldg(pr + n + p) - tl.y, d2 = __ldg(pr + 2 * n + p) - tl.z;
            float e = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float d = __ldg(pr + (size_t)(3 + k) * n + p) - qs[k];
                e += (d * d) * mr;
            }
            acc[l * 4 + 0] += (double)(((d0 * d0) * mt + (d1 * d1) * mt) + (d2 * d2) * mt);
            acc[l * 4 + 1] += (double)e;
            acc[l * 4 + 2] += (double)mt;
            acc[l * 4 + 3] += (double)mr;
        }
    }
    cluster_sum<12>(acc, sh, xch_d);
    if (tid == 0 && blockIdx.x == 0) {
        float* sv = save + (size_t)b * LT_SAVE;
        float lt = 0.f, lr = 0.f;
        for (int k = 0; k < 3; ++k) {
            const float d = T_pred[b * 3 + k] - s_lab[k];
            lt += d * d;
        }
        for (int k = 0; k < 4; ++k) {
            const float d = q_pred[b * 4 + k] - s_lab[3 + k];
            lr += d * d;
        }
        sv[0] = lt / (3.f + 1e-12f);
        sv[1] = lr / (4.f + 1e-12f);
        for (int l = 0; l < 3; ++l) {
            const float mt = (float)acc[l * 4 + 2], mr = (float)acc[l * 4 + 3];
            sv[2 + l] = (float)acc[l * 4 + 0] / (3.f * mt + 1e-12f);
            sv[5 + l] = (float)acc[l * 4 + 1] / (4.f * mr + 1e-12f);
            sv[8 + l] = mt;
            sv[11 + l] = mr;
        }
        for (int k = 0; k < 7; ++k) sv[14 + k] = s_lab[k];
        __threadfence();
        s_last = atomicAdd(counter, 1) == B - 1;
    }
    __syncthreads();
    if (s_last && tid < 8) {
        __threadfence();
        // loss = w * ( sum_b e^{-alpha} L_b / (B + 1e-12) + alpha )      (focal_gamma 0: uniform weights)
        const float a = tid == 0 ? *alpha_t : (tid == 1 ? *alpha_r : (tid < 5 ? *alpha_pt : *alpha_pr));
        const float w = tid == 0 ? w_t : (tid == 1 ? w_r : (tid < 5 ? w_pt : w_pr));
        const float fw = 1.f / ((float)B + 1e-12f);
        float sum = 0.f;
        for (int bb = 0; bb < B; ++bb) sum += fw * (expf(-a) * __ldcg(save + (size_t)bb * LT_SAVE + tid));
        losses[tid] = w * (sum + a);
        if (tid == 0) *counter = 0;
    }
}

// gradients: dT [B,3], dq [B,4], d pyramid preds [B,7,h,w] x3, dalpha [4] (t, r, pyramid t, pyramid r)
__global__ void __launch_bounds__(PT_THREADS)
k_loss_tail_bwd(const float* __restrict__ T_pred, const float* __restrict__ q_pred, PyrLevel L0, PyrLevel L1, PyrLevel L2,
                rslo_tq_geom_t G, int B, const float* __restrict__ alpha_t, const float* __restrict__ alpha_r,
                const float* __restrict__ alpha_pt, const float* __restrict__ alpha_pr, float w_t, float w_r, float w_pt,
                float w_pr, const float* __restrict__ save, const float* __restrict__ g_losses /* [8] */,
                float* __restrict__ dT, float* __restrict__ dq, float* __restrict__ dalpha /* [4], zeroed */)
{
    const int b = blockIdx.y, tid = threadIdx.x;         // grid (chunks, B)
    const float* sv = save + (size_t)b * LT_SAVE;
    const float fw = 1.f / ((float)B + 1e-12f);
    const V3 ts = V3{sv[14], sv[15], sv[16]};
    const float qs[4] = {sv[17], sv[18], sv[19], sv[20]};
    const float qi[4] = {qs[0], -qs[1], -qs[2], -qs[3]};
    const float eat = expf(-*alpha_t), ear = expf(-*alpha_r), eapt = expf(-*alpha_pt), eapr = expf(-*alpha_pr);
    if (tid == 0 && blockIdx.x == 0) {
        const float ct = g_losses[0] * w_t * fw * eat * 2.f / (3.f + 1e-12f);
        const float cr = g_losses[1] * w_r * fw * ear * 2.f / (4.f + 1e-12f);
        for (int k = 0; k < 3; ++k) dT[b * 3 + k] = ct * (T_pred[b * 3 + k] - sv[14 + k]);
        for (int k = 0; k < 4; ++k) dq[b * 4 + k] = cr * (q_pred[b * 4 + k] - sv[17 + k]);
        // d/dalpha [ w (sum_b fw e^{-a} L_b + a) ] = w (1 [once] - fw e^{-a} L_b [per pair])
        float da[4] = {-g_losses[0] * w_t * fw * eat * sv[0], -g_losses[1] * w_r * fw * ear * sv[1], 0.f, 0.f};
        for (int l = 0; l < 3; ++l) {
            da[2] -= g_losses[2 + l] * w_pt * fw * eapt * sv[2 + l];
            da[3] -= g_losses[5 + l] * w_pr * fw * eapr * sv[5 + l];
        }
        if (b == 0) {
            da[0] += g_losses[0] * w_t;
            da[1] += g_losses[1] * w_r;
            for (int l = 0; l < 3; ++l) {
                da[2] += g_losses[2 + l] * w_pt;
                da[3] += g_losses[5 + l] * w_pr;
            }
        }
        for (int k = 0; k < 4; ++k) atomicAdd(dalpha + k, da[k]);
    }
    const PyrLevel Ls[3] = {L0, L1, L2};
#pragma unroll
    for (int l = 0; l < 3; ++l) {
        const PyrLevel& L = Ls[l];
        const int n = L.h * L.w;
        const float* pr = L.pred + (size_t)b * 7 * n;
        const float* mk = L.mask + (size_t)b * 2 * n;
        float* d = L.dpred + (size_t)b * 7 * n;
        const float ct = g_losses[2 + l] * w_pt * fw * eapt * 2.f / (3.f * sv[8 + l] + 1e-12f);
        const float cr = g_losses[5 + l] * w_pr * fw * eapr * 2.f / (4.f * sv[11 + l] + 1e-12f);
        for (int p = blockIdx.x * PT_THREADS + tid; p < n; p += gridDim.x * PT_THREADS) {
            const int i = p / L.w, j = p - i * L.w;
            V3 tl;
            target_at(G, i * L.stride, j * L.stride, ts, qi, tl);
            const float mt = __ldg(mk + p) * ct, mr = __ldg(mk + n + p) * cr;
            d[p] = mt * (__ldg(pr + p) - tl.x);
            d[n + p] = mt * (__ldg(pr + n + p) - tl.y);
            d[2 * n + p] = mt * (__ldg(pr + 2 * n + p) - tl.z);
#pragma unroll
            for (int k = 0; k < 4; ++k) d[(size_t)(3 + k) * n + p] = mr * (__ldg(pr + (size_t)(3 + k) * n + p) - qs[k]);
        }
    }
}

}  // namespace
}  // namespace rslo

using namespace rslo;

extern "C" int rslo_head_tail_forward(const float* tq32, const float* t_logit32, const float* r_logit32, const float* mask,
                                      const float* py0_32, const float* py1_32, int B, rslo_tq_geom_t geom, float* pose_t,
                                      float* pose_q, float* tq_map_g, float* t_conf, float* r_conf, float* pm2_pred,
                                      float* pm2_mask, float* pm1_pred, float* pm1_mask, float* pm0_pred, float* pm0_mask,
                                      float* occ1, float* occ0, float* save, rslo_stream_t stream)
{
    if (B <= 0) return 0;
    if (geom.H % 4 || geom.W % 4) {
        set_last_error("rslo_head_tail_forward: H, W must be multiples of 4", cudaErrorInvalidValue);
        return (int)cudaErrorInvalidValue;
    }
    RSLO_COUNT();
    k_head_tail_fwd<<<dim3(PT_CLUSTER, B), PT_THREADS, 0, (cudaStream_t)stream>>>(tq32, t_logit32, r_logit32, mask, py0_32, py1_32, geom, pose_t,
                                                               pose_q, tq_map_g, t_conf, r_conf, pm2_pred, pm2_mask, pm1_pred,
                                                               pm1_mask, pm0_pred, pm0_mask, occ1, occ0, save);
    RSLO_CHECK_LAUNCH("rslo_head_tail_forward");
    return 0;
}

extern "C" int rslo_head_tail_backward(const float* tq32, const float* mask, const float* t_conf, const float* r_conf,
                                       const float* occ1, const float* occ0, const float* save, int B, rslo_tq_geom_t geom,
                                       const float* g_pose_t, const float* g_pose_q, const float* g_tq_map_g,
                                       const float* g_t_conf, const float* g_r_conf, const float* g_pm2_pred,
                                       const float* g_pm1_pred, const float* g_pm0_pred, float* d_tq32, float* d_t_logit32,
                                       float* d_r_logit32, float* d_py1_32, float* d_py0_32, rslo_stream_t stream)
{
    if (B <= 0) return 0;
    RSLO_COUNT();
    int chunks = 148 / B;
    if (chunks < 1) chunks = 1;
    if (chunks > cdiv(geom.H * geom.W, PT_THREADS)) chunks = cdiv(geom.H * geom.W, PT_THREADS);
    k_head_tail_bwd<<<dim3(chunks, B), PT_THREADS, 0, (cudaStream_t)stream>>>(tq32, mask, t_conf, r_conf, occ1, occ0, save, geom, g_pose_t,
                                                               g_pose_q, g_tq_map_g, g_t_conf, g_r_conf, g_pm2_pred, g_pm1_pred,
                                                               g_pm0_pred, d_tq32, d_t_logit32, d_r_logit32, d_py1_32, d_py0_32);
    RSLO_CHECK_LAUNCH("rslo_head_tail_backward");
    return 0;
}

static PyrLevel make_level(const float* pred, const float* mask, float* dpred, int H, int W, int stride)
{
    PyrLevel L;
    L.pred = pred; L.mask = mask; L.dpred = dpred; L.h = H / stride; L.w = W / stride; L.stride = stride;
    return L;
}

extern "C" int rslo_loss_tail_forward(const float* T_pred, const float* q_pred, const float* pm0_pred, const float* pm0_mask,
                                      const float* pm1_pred, const float* pm1_mask, const float* pm2_pred,
                                      const float* pm2_mask, const float* res_r, const float* res_t, int identity_pose, int B,
                                      rslo_tq_geom_t geom, const float* alpha_t, const float* alpha_r,
                                      const float* alpha_pt, const float* alpha_pr, float w_t, float w_r, float w_pt, float w_pr,
                                      float* tq_target, float* save, float* losses8, int* counter, rslo_stream_t stream)
{
    if (B <= 0) return 0;
    RSLO_COUNT();
    k_loss_tail_fwd<<<dim3(PT_CLUSTER, B), PT_THREADS, 0, (cudaStream_t)stream>>>(
        T_pred, q_pred, make_level(pm0_pred, pm0_mask, nullptr, geom.H, geom.W, 4),
        make_level(pm1_pred, pm1_mask, nullptr, geom.H, geom.W, 2), make_level(pm2_pred, pm2_mask, nullptr, geom.H, geom.W, 1),
        res_r, res_t, identity_pose, geom, B, alpha_t, alpha_r, alpha_pt, alpha_pr, w_t, w_r, w_pt, w_pr, tq_target, save,
        losses8, counter);
    RSLO_CHECK_LAUNCH("rslo_loss_tail_forward");
    return 0;
}

extern "C" int rslo_loss_tail_backward(const float* T_pred, const float* q_pred, const float* pm0_pred, const float* pm0_mask,
                                       const float* pm1_pred, const float* pm1_mask, const float* pm2_pred,
                                       const float* pm2_mask, int B, rslo_tq_geom_t geom, const float* alpha_t,
                                       const float* alpha_r, const float* alpha_pt, const float* alpha_pr, float w_t, float w_r,
                                       float w_pt, float w_pr, const float* save, const float* g_losses8, float* dT, float* dq,
                                       float* d_pm0, float* d_pm1, float* d_pm2, float* dalpha4, rslo_stream_t stream)
{
    if (B <= 0) return 0;
    RSLO_CHECK(cudaMemsetAsync(dalpha4, 0, 4 * sizeof(float), (cudaStream_t)stream));
    RSLO_COUNT();
    int chunks = 148 / B;
    if (chunks < 1) chunks = 1;
    if (chunks > cdiv(geom.H * geom.W, PT_THREADS)) chunks = cdiv(geom.H * geom.W, PT_THREADS);
    k_loss_tail_bwd<<<dim3(chunks, B), PT_THREADS, 0, (cudaStream_t)stream>>>(
        T_pred, q_pred, make_level(pm0_pred, pm0_mask, d_pm0, geom.H, geom.W, 4),
        make_level(pm1_pred, pm1_mask, d_pm1, geom.H, geom.W, 2), make_level(pm2_pred, pm2_mask, d_pm2, geom.H, geom.W, 1), geom,
        B, alpha_t, alpha_r, alpha_pt, alpha_pr, w_t, w_r, w_pt, w_pr, save, g_losses8, dT, dq, dalpha4);
    RSLO_CHECK_LAUNCH("rslo_loss_tail_backward");
    return 0;
}
