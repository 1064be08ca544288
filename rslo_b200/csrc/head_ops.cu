// Elementwise / reduction companions of the dense tensor-core convolutions (conv2d_tc.cu): everything the
// odometry head does BETWEEN its convolutions, in the layout the convolutions consume (NHWC rows, stored as the
// split pair hi = RN_tf32(x), lo = RN_tf32(x - hi)) so no activation is ever re-laid-out or re-split by a separate pass.
//
//   k_head_pack / k_head_unpack : NCHW BEV maps of a frame pair <-> NHWC split pair (+ first-frame occupancy mask,
//                                 `rslo/models/odom_pred.py:165-168`)
//   k_bn_act_fwd                : BatchNorm (batch statistics accumulated by the producing convolution's epilogue, or
//                                 running statistics) + residual add + ReLU -> fp32 activation + its split pair;
//                                 running-statistics update (`odom_pred_base.py:140-141`: eps 1e-3, momentum 0.01;
//                                 `custom_resnet_spc.py:224-298` residual blocks)
//   k_bn_bwd_reduce / _apply    : its backward (ReLU mask, sum(dy), sum(dy*xhat), input gradient written as the split
//                                 pair the data- and weight-gradient convolutions read, residual gradient, dgamma/dbeta)
//   k_upcat_split / k_upcat_bwd : nearest-neighbour upsampling + channel concat of the decoder (`odom_pred_base.py:196-207`)
//   k_bias_grad                 : bias gradient of the narrow output convolutions
//   k_multi_prepare / _finish   : weight images of every convolution of the head in one launch; weight gradients of every
//                                 convolution back to OIHW in one launch
// All HBM/L2-bound streaming kernels: float4 accesses, one pass per tensor.
#include "tc_common.cuh"

namespace rslo {
namespace {
using tc::tf32_rn;

constexpr int HE_THREADS = 256;

__device__ __forceinline__ void split4(const float4& v, float4& h, float4& l)
{
    h = make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
    l = make_float4(tf32_rn(v.x - h.x), tf32_rn(v.y - h.y), tf32_rn(v.z - h.z), tf32_rn(v.w - h.w));
}

// ---------------------------------------------------------------------------------------------------
// NCHW pair -> NHWC split pair.  grid (ceil(HW/32), B), block (32, 8)
// ---------------------------------------------------------------------------------------------------
__global__ void k_head_pack(const float* __restrict__ x1, const float* __restrict__ x2, int C, int HW, float* __restrict__ hi,
                            float* __restrict__ lo, float* __restrict__ mask)
{
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float tile[32][33];
    __shared__ float psum[8][32];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int p0 = blockIdx.x * 32, b = blockIdx.y;
    const int C2 = 2 * C;
    float s = 0.f;
    for (int ct = 0; ct < C2 / 32; ++ct) {
        const bool first = ct * 32 < C;
        const float* src = first ? x1 + ((size_t)b * C + ct * 32) * HW : x2 + ((size_t)b * C + ct * 32 - C) * HW;
        for (int j = ty; j < 32; j += 8) {
            const int p = p0 + tx;
            const float v = p < HW ? __ldg(src + (size_t)j * HW + p) : 0.f;
            tile[j][tx] = v;
            if (first) s += v;
        }
        __syncthreads();
        for (int j = ty; j < 32; j += 8) {
            const int p = p0 + j;
            if (p < HW) {
                const float v = tile[tx][j];
                const float h = tf32_rn(v);
                const size_t o = ((size_t)b * HW + p) * C2 + ct * 32 + tx;
                hi[o] = h;
                lo[o] = tf32_rn(v - h);
            }
        }
        __syncthreads();
    }
    psum[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && p0 + tx < HW) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += psum[k][tx];
        mask[(size_t)b * HW + p0 + tx] = t != 0.f ? 1.f : 0.f;
    }
}

// NHWC gradient [B][HW][2C] -> NCHW gradients of the two frames.  grid (ceil(HW/32), 2C/32, B), block (32, 8)
__global__ void k_head_unpack(const float* __restrict__ dx, int C, int HW, float* __restrict__ g1, float* __restrict__ g2)
{
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float tile[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int p0 = blockIdx.x * 32, ct = blockIdx.y, b = blockIdx.z;
    const int C2 = 2 * C;
    for (int j = ty; j < 32; j += 8) {
        const int p = p0 + j;
        tile[j][tx] = p < HW ? __ldg(dx + ((size_t)b * HW + p) * C2 + ct * 32 + tx) : 0.f;
    }
    __syncthreads();
    float* dst = ct * 32 < C ? g1 + ((size_t)b * C + ct * 32) * HW : g2 + ((size_t)b * C + ct * 32 - C) * HW;
    for (int j = ty; j < 32; j += 8) {
        const int p = p0 + tx;
        if (p < HW) dst[(size_t)j * HW + p] = tile[tx][j];
    }
}

// ---------------------------------------------------------------------------------------------------
// BatchNorm + residual + ReLU forward.  grid (chunks, B), block 256, smem 3*C floats
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(HE_THREADS)
k_bn_act_fwd(const float* __restrict__ y, int HW, int C, int ipg, int G, int chunk, const double* __restrict__ stats,
             const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ running_mean,
             float* __restrict__ running_var, long long* __restrict__ nbt, float eps, float momentum, int update_repeat,
             const float* __restrict__ residual, int relu, float* __restrict__ z, float* __restrict__ zhi,
             float* __restrict__ zlo, float* __restrict__ mean_rstd)
{
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float sm[];
    float* s_mean = sm;
    float* s_rs = sm + C;
    float* s_beta = sm + 2 * C;
    const int b = blockIdx.y, g = b / ipg;
    const double n = (double)ipg * HW;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float mean, rstd;
        if (stats != nullptr) {
            const double m = stats[((size_t)g * C + c) * 2] / n;
            double var = stats[((size_t)g * C + c) * 2 + 1] / n - m * m;
            if (var < 0) var = 0;
            mean = (float)m;
            rstd = (float)(1.0 / sqrt(var + (double)eps));
        } else {
            mean = running_mean[c];
            rstd = 1.0f / sqrtf(running_var[c] + eps);
        }
        s_mean[c] = mean;
        s_rs[c] = rstd * __ldg(gamma + c);
        s_beta[c] = __ldg(beta + c);
        if (mean_rstd != nullptr && blockIdx.x == 0 && b % ipg == 0) {
            mean_rstd[((size_t)g * C + c) * 2] = mean;
            mean_rstd[((size_t)g * C + c) * 2 + 1] = rstd;
        }
    }
    // running statistics: one block walks the statistics groups in order (= the samples of the step seen one
    // forward call after the other); `update_repeat` = how often the reference runs this layer per forward
    if (stats != nullptr && running_mean != nullptr && blockIdx.x == 0 && b == 0 && update_repeat > 0) {
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            float rm = running_mean[c], rv = running_var[c];
            for (int gg = 0; gg < G; ++gg) {
                const double m = stats[((size_t)gg * C + c) * 2] / n;
                double var = stats[((size_t)gg * C + c) * 2 + 1] / n - m * m;
                if (var < 0) var = 0;
                const float mu = (float)m, vu = (float)(var * (n / (n - 1.0)));
                for (int r = 0; r < update_repeat; ++r) {
                    rm = (1.f - momentum) * rm + momentum * mu;
                    rv = (1.f - momentum) * rv + momentum * vu;
                }
            }
            running_mean[c] = rm;
            running_var[c] = rv;
        }
        if (threadIdx.x == 0 && nbt != nullptr) *nbt += (long long)G * update_repeat;
    }
    __syncthreads();
    const int c4n = C >> 2;
    const int p0 = blockIdx.x * chunk;
    const int p1 = min(HW, p0 + chunk);
    const int items = (p1 - p0) * c4n;
    const size_t base = ((size_t)b * HW + p0) * c4n;
    const float4* y4 = reinterpret_cast<const float4*>(y) + base;
    const float4* r4 = residual ? reinterpret_cast<const float4*>(residual) + base : nullptr;
    float4* z4 = z ? reinterpret_cast<float4*>(z) + base : nullptr;
    float4* h4 = zhi ? reinterpret_cast<float4*>(zhi) + base : nullptr;
    float4* l4 = zlo ? reinterpret_cast<float4*>(zlo) + base : nullptr;
    for (int i = threadIdx.x; i < items; i += blockDim.x) {
        const int c = (i % c4n) * 4;
        const float4 v = __ldg(y4 + i);
        float4 o;
        o.x = (v.x - s_mean[c]) * s_rs[c] + s_beta[c];
        o.y = (v.y - s_mean[c + 1]) * s_rs[c + 1] + s_beta[c + 1];
        o.z = (v.z - s_mean[c + 2]) * s_rs[c + 2] + s_beta[c + 2];
        o.w = (v.w - s_mean[c + 3]) * s_rs[c + 3] + s_beta[c + 3];
        if (r4) {
            const float4 r = __ldg(r4 + i);
            o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        }
        if (relu) {
            o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
        }
        if (z4) z4[i] = o;
        if (h4) {
            float4 h, l;
            split4(o, h, l);
            h4[i] = h;
            l4[i] = l;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// backward, pass 1: sums[g][c] += { sum(dyh), sum(dyh * xhat) },  dyh = dz * [z > 0]
// grid (chunks, B), block 256 (C/4 must divide 256)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(HE_THREADS)
k_bn_bwd_reduce(const float* __restrict__ dz, const float* __restrict__ z, const float* __restrict__ y, int HW, int C, int ipg,
                int chunk, const float* __restrict__ mean_rstd, int relu, double* __restrict__ sums)
{
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float4 red[2][HE_THREADS];
    const int b = blockIdx.y, g = b / ipg;
    const int c4n = C >> 2;
    const int lanes = HE_THREADS / c4n;
    const int c4 = threadIdx.x % c4n, lane = threadIdx.x / c4n;
    const int c = c4 * 4;
    const float* mr = mean_rstd + ((size_t)g * C + c) * 2;
    const float m0 = mr[0], r0 = mr[1], m1 = mr[2], r1 = mr[3], m2 = mr[4], r2 = mr[5], m3 = mr[6], r3 = mr[7];
    const int p0 = blockIdx.x * chunk;
    const int p1 = min(HW, p0 + chunk);
    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
    for (int p = p0 + lane; p < p1; p += lanes) {
        const size_t i = ((size_t)b * HW + p) * c4n + c4;
        float4 d = __ldg(reinterpret_cast<const float4*>(dz) + i);
        if (relu) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(z) + i);
            d.x = a.x > 0.f ? d.x : 0.f; d.y = a.y > 0.f ? d.y : 0.f; d.z = a.z > 0.f ? d.z : 0.f; d.w = a.w > 0.f ? d.w : 0.f;
        }
        const float4 v = __ldg(reinterpret_cast<const float4*>(y) + i);
        s1.x += d.x; s1.y += d.y; s1.z += d.z; s1.w += d.w;
        s2.x = fmaf(d.x, (v.x - m0) * r0, s2.x);
        s2.y = fmaf(d.y, (v.y - m1) * r1, s2.y);
        s2.z = fmaf(d.z, (v.z - m2) * r2, s2.z);
        s2.w = fmaf(d.w, (v.w - m3) * r3, s2.w);
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
        double* dst = sums + ((size_t)g * C + c) * 2;
        atomicAdd(dst + 0, (double)s1.x); atomicAdd(dst + 1, (double)s2.x);
        atomicAdd(dst + 2, (double)s1.y); atomicAdd(dst + 3, (double)s2.y);
        atomicAdd(dst + 4, (double)s1.z); atomicAdd(dst + 5, (double)s2.z);
        atomicAdd(dst + 6, (double)s1.w); atomicAdd(dst + 7, (double)s2.w);
    }
}

// backward, pass 2: dy = gamma*rstd * (dyh - mean(dyh) - xhat*mean(dyh*xhat)) as a split pair; residual gradient; dgamma/dbeta
// grid (chunks, B), block 256, smem 5*C floats
__global__ void __launch_bounds__(HE_THREADS)
k_bn_bwd_apply(const float* __restrict__ dz, const float* __restrict__ z, const float* __restrict__ y, int HW, int C, int ipg,
               int G, int chunk, const float* __restrict__ mean_rstd, const float* __restrict__ gamma,
               const double* __restrict__ sums, int relu, int batch_stats, float* __restrict__ ghi, float* __restrict__ glo,
               float* __restrict__ dres, int dres_add, float* __restrict__ dgamma, float* __restrict__ dbeta,
               float* __restrict__ dbias)
{
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float sm[];
    float* s_mean = sm;
    float* s_rstd = sm + C;
    float* s_a = sm + 2 * C;
    float* s_m1 = sm + 3 * C;
    float* s_m2 = sm + 4 * C;
    const int b = blockIdx.y, g = b / ipg;
    const double n = (double)ipg * HW;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float mean = mean_rstd[((size_t)g * C + c) * 2], rstd = mean_rstd[((size_t)g * C + c) * 2 + 1];
        s_mean[c] = mean;
        s_rstd[c] = rstd;
        s_a[c] = __ldg(gamma + c) * rstd;
        s_m1[c] = batch_stats ? (float)(sums[((size_t)g * C + c) * 2] / n) : 0.f;
        s_m2[c] = batch_stats ? (float)(sums[((size_t)g * C + c) * 2 + 1] / n) : 0.f;
    }
    if (blockIdx.x == 0 && b == 0) {
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            double sb = 0, sg = 0;
            for (int gg = 0; gg < G; ++gg) {
                sb += sums[((size_t)gg * C + c) * 2];
                sg += sums[((size_t)gg * C + c) * 2 + 1];
            }
            if (dbeta) dbeta[c] = (float)sb;
            if (dgamma) dgamma[c] = (float)sg;
            // bias of the producing convolution: sum(dy) = 0 under batch statistics, gamma*rstd*sum(dyh) otherwise
            if (dbias) {
                double t = 0;
                if (!batch_stats)
                    for (int gg = 0; gg < G; ++gg)
                        t += sums[((size_t)gg * C + c) * 2] * (double)(__ldg(gamma + c) * mean_rstd[((size_t)gg * C + c) * 2 + 1]);
                dbias[c] = (float)t;
            }
        }
    }
    __syncthreads();
    const int c4n = C >> 2;
    const int p0 = blockIdx.x * chunk;
    const int p1 = min(HW, p0 + chunk);
    const int items = (p1 - p0) * c4n;
    const size_t base = ((size_t)b * HW + p0) * c4n;
    const float4* d4 = reinterpret_cast<const float4*>(dz) + base;
    const float4* z4 = reinterpret_cast<const float4*>(z) + base;
    const float4* y4 = reinterpret_cast<const float4*>(y) + base;
    float4* h4 = reinterpret_cast<float4*>(ghi) + base;
    float4* l4 = reinterpret_cast<float4*>(glo) + base;
    float4* r4 = dres ? reinterpret_cast<float4*>(dres) + base : nullptr;
    for (int i = threadIdx.x; i < items; i += blockDim.x) {
        const int c = (i % c4n) * 4;
        float4 d = __ldg(d4 + i);
        if (relu) {
            const float4 a = __ldg(z4 + i);
            d.x = a.x > 0.f ? d.x : 0.f; d.y = a.y > 0.f ? d.y : 0.f; d.z = a.z > 0.f ? d.z : 0.f; d.w = a.w > 0.f ? d.w : 0.f;
        }
        if (r4) {
            float4 r = d;
            if (dres_add) {
                const float4 o = r4[i];
                r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w;
            }
            r4[i] = r;
        }
        const float4 v = __ldg(y4 + i);
        float4 o;
        o.x = s_a[c] * (d.x - s_m1[c] - (v.x - s_mean[c]) * s_rstd[c] * s_m2[c]);
        o.y = s_a[c + 1] * (d.y - s_m1[c + 1] - (v.y - s_mean[c + 1]) * s_rstd[c + 1] * s_m2[c + 1]);
        o.z = s_a[c + 2] * (d.z - s_m1[c + 2] - (v.z - s_mean[c + 2]) * s_rstd[c + 2] * s_m2[c + 2]);
        o.w = s_a[c + 3] * (d.w - s_m1[c + 3] - (v.w - s_mean[c + 3]) * s_rstd[c + 3] * s_m2[c + 3]);
        float4 h, l;
        split4(o, h, l);
        h4[i] = h;
        l4[i] = l;
    }
}

// ---------------------------------------------------------------------------------------------------
// nearest upsample (factor u) + channel concat: z [B][H][W][C] -> split planes [B][uH][uW][ld] at channel `choff`
// ---------------------------------------------------------------------------------------------------
__global__ void k_upcat_split(const float4* __restrict__ z, int B, int H, int W, int c4n, int u, int ld4, int choff4,
                              float4* __restrict__ hi, float4* __restrict__ lo)
{
    pdl_launch_dependents();
    pdl_wait();
    const size_t total = (size_t)B * H * W * c4n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % c4n);
        size_t p = i / c4n;
        const int w = (int)(p % W);
        p /= W;
        const int h = (int)(p % H), b = (int)(p / H);
        float4 hv, lv;
        split4(__ldg(z + i), hv, lv);
        for (int dy = 0; dy < u; ++dy)
            for (int dx = 0; dx < u; ++dx) {
                const size_t o = (((size_t)b * H * u + (size_t)(h * u + dy)) * W * u + (size_t)(w * u + dx)) * ld4 + choff4 + c4;
                hi[o] = hv;
                lo[o] = lv;
            }
    }
}

__global__ void k_upcat_bwd(const float4* __restrict__ dcat, int B, int H, int W, int c4n, int u, int ld4, int choff4,
                            float4* __restrict__ dz, int add)
{
    pdl_launch_dependents();
    pdl_wait();
    const size_t total = (size_t)B * H * W * c4n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % c4n);
        size_t p = i / c4n;
        const int w = (int)(p % W);
        p /= W;
        const int h = (int)(p % H), b = (int)(p / H);
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int dy = 0; dy < u; ++dy)
            for (int dx = 0; dx < u; ++dx) {
                const float4 v = __ldg(dcat + (((size_t)b * H * u + (size_t)(h * u + dy)) * W * u + (size_t)(w * u + dx)) * ld4 +
                                       choff4 + c4);
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
        if (add) {
            const float4 o = dz[i];
            s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
        }
        dz[i] = s;
    }
}

// bias gradient of a narrow output convolution: out[c] += sum_rows g[row][c], c < C <= 32 (out pre-zeroed)
__global__ void k_bias_grad(const float* __restrict__ g, int N, int ld, int C, float* __restrict__ out)
{
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float red[8][32];
    const int c = threadIdx.x & 31, l = threadIdx.x >> 5;
    float s = 0.f;
    if (c < C)
        for (int r = blockIdx.x * 8 + l; r < N; r += gridDim.x * 8) s += __ldg(g + (size_t)r * ld + c);
    red[l][c] = s;
    __syncthreads();
    if (l == 0 && c < C) {
#pragma unroll
        for (int k = 1; k < 8; ++k) s += red[k][c];
        atomicAdd(out + c, s);
    }
}

// ---------------------------------------------------------------------------------------------------
// table-driven weight preparation / weight-gradient finish: one launch for all convolutions of the head
// ---------------------------------------------------------------------------------------------------
// grid (blocks, n_layers)
__global__ void k_multi_prepare(const rslo_conv_prep_t* __restrict__ tab)
{
    pdl_launch_dependents();
    pdl_wait();
    const rslo_conv_prep_t e = tab[blockIdx.y];
    const int total = e.taps * e.CoutP * e.Cin;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int t = i / (e.CoutP * e.Cin), r = i - t * e.CoutP * e.Cin;
        if (e.img_fwd) {                        // [t][co][ci]
            const int co = r / e.Cin, ci = r - co * e.Cin;
            const float v = co < e.Cout ? __ldg(e.w + ((size_t)co * e.Cin + ci) * e.taps + t) : 0.f;
            const float h = tf32_rn(v);
            e.img_fwd[i] = h;
            e.img_fwd[(size_t)total + i] = tf32_rn(v - h);
        }
        if (e.img_bwd) {                        // [t][ci][co]
            const int ci = r / e.CoutP, co = r - ci * e.CoutP;
            const float v = co < e.Cout ? __ldg(e.w + ((size_t)co * e.Cin + ci) * e.taps + t) : 0.f;
            const float h = tf32_rn(v);
            e.img_bwd[i] = h;
            e.img_bwd[(size_t)total + i] = tf32_rn(v - h);
        }
    }
    if (e.bias_pad && blockIdx.x == 0)
        for (int c = threadIdx.x; c < e.CoutP; c += blockDim.x) e.bias_pad[c] = (e.bias && c < e.Cout) ? __ldg(e.bias + c) : 0.f;
}

// dW [taps][Cin][CoutP] -> OIHW gradient [Cout][Cin][taps].  grid (blocks, n_layers)
__global__ void k_multi_wgrad_finish(const rslo_wgrad_finish_t* __restrict__ tab)
{
    pdl_launch_dependents();
    pdl_wait();
    const rslo_wgrad_finish_t e = tab[blockIdx.y];
    const int total = e.taps * e.Cout * e.Cin;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int t = i % e.taps, ci = (i / e.taps) % e.Cin, co = i / (e.taps * e.Cin);
        e.gw[i] = __ldg(e.dW + ((size_t)t * e.Cin + ci) * e.CoutP + co);
    }
}

static int pick_chunk(int HW, int c4n, int B)
{
    // ~512 float4 items per block (2 per thread: the passes are latency-bound streaming loops, r02e_head_bn_ncu.md showed
    // 12 % warps active at 2048), at least ~4 blocks per SM over the whole launch when the tensor allows
    long long items = (long long)HW * c4n;
    int chunks = (int)((items + 511) / 512);
    const int want = (592 + B - 1) / B;
    if (chunks < want && items >= 256LL * want) chunks = want;
    if (chunks < 1) chunks = 1;
    int chunk = (HW + chunks - 1) / chunks;
    if (chunk < 1) chunk = 1;
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

extern "C" int rslo_head_pack_input(const float* x1, const float* x2, int B, int C, int HW, float* split_pair, float* mask,
                                    rslo_stream_t stream)
{
    if (C % 32 != 0) return bad("rslo_head_pack_input: C must be a multiple of 32");
    const size_t plane = (size_t)B * HW * 2 * C;
    RSLO_COUNT();
    launch_pdl(k_head_pack, dim3(cdiv(HW, 32), B), dim3(32, 8), 0, (cudaStream_t)stream, x1, x2, C, HW, split_pair, split_pair + plane, mask);
    RSLO_CHECK_LAUNCH("rslo_head_pack_input");
    return 0;
}

extern "C" int rslo_head_unpack_grad(const float* dx, int B, int C, int HW, float* g1, float* g2, rslo_stream_t stream)
{
    if (C % 32 != 0) return bad("rslo_head_unpack_grad: C must be a multiple of 32");
    RSLO_COUNT();
    launch_pdl(k_head_unpack, dim3(cdiv(HW, 32), 2 * C / 32, B), dim3(32, 8), 0, (cudaStream_t)stream, dx, C, HW, g1, g2);
    RSLO_CHECK_LAUNCH("rslo_head_unpack_grad");
    return 0;
}

extern "C" int rslo_bn_act_forward(const float* y, int B, int HW, int C, int imgs_per_group, const double* stats,
                                   const float* gamma, const float* beta, float* running_mean, float* running_var,
                                   long long* num_batches_tracked, float eps, float momentum, int update_repeat,
                                   const float* residual, int relu, float* z, float* z_split, float* mean_rstd,
                                   rslo_stream_t stream)
{
    if (C % 4 != 0 || imgs_per_group < 1 || B % imgs_per_group != 0) return bad("rslo_bn_act_forward: bad shape");
    if (stats == nullptr && running_mean == nullptr) return bad("rslo_bn_act_forward: no statistics");
    const int chunk = pick_chunk(HW, C / 4, B);
    const size_t plane = (size_t)B * HW * C;
    RSLO_COUNT();
    launch_pdl(k_bn_act_fwd, dim3(cdiv(HW, chunk), B), HE_THREADS, 3 * C * sizeof(float), (cudaStream_t)stream, 
        y, HW, C, imgs_per_group, B / imgs_per_group, chunk, stats, gamma, beta, running_mean, running_var, num_batches_tracked,
        eps, momentum, update_repeat, residual, relu, z, z_split, z_split ? z_split + plane : nullptr, mean_rstd);
    RSLO_CHECK_LAUNCH("rslo_bn_act_forward");
    return 0;
}

extern "C" int rslo_bn_act_backward(const float* dz, const float* z, const float* y, int B, int HW, int C, int imgs_per_group,
                                    const float* mean_rstd, const float* gamma, int relu, int batch_stats, double* sums,
                                    float* g_split, float* dres, int dres_accumulate, float* dgamma, float* dbeta,
                                    float* dbias, rslo_stream_t stream)
{
    const int c4n = C / 4;
    if (C % 4 != 0 || c4n > HE_THREADS || HE_THREADS % c4n != 0 || imgs_per_group < 1 || B % imgs_per_group != 0)
        return bad("rslo_bn_act_backward: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    const int chunk = pick_chunk(HW, c4n, B);
    const size_t plane = (size_t)B * HW * C;
    RSLO_COUNT();
    launch_pdl(k_bn_bwd_reduce, dim3(cdiv(HW, chunk), B), HE_THREADS, 0, st, dz, z, y, HW, C, imgs_per_group, chunk, mean_rstd, relu, sums);
    RSLO_CHECK_LAUNCH("rslo_bn_act_backward(reduce)");
    RSLO_COUNT();
    launch_pdl(k_bn_bwd_apply, dim3(cdiv(HW, chunk), B), HE_THREADS, 5 * C * sizeof(float), st, 
        dz, z, y, HW, C, imgs_per_group, B / imgs_per_group, chunk, mean_rstd, gamma, sums, relu, batch_stats, g_split,
        g_split + plane, dres, dres_accumulate, dgamma, dbeta, dbias);
    RSLO_CHECK_LAUNCH("rslo_bn_act_backward(apply)");
    return 0;
}

extern "C" int rslo_upcat_split(const float* z, int B, int H, int W, int C, int up, int ld, int choff, float* dst_split,
                                rslo_stream_t stream)
{
    if (C % 4 || ld % 4 || choff % 4 || up < 1) return bad("rslo_upcat_split: bad shape");
    const size_t total = (size_t)B * H * W * (C / 4);
    const size_t plane = (size_t)B * H * up * W * up * ld;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    RSLO_COUNT();
    launch_pdl(k_upcat_split, blocks, 256, 0, (cudaStream_t)stream, (const float4*)z, B, H, W, C / 4, up, ld / 4, choff / 4,
                                                           (float4*)dst_split, (float4*)(dst_split + plane));
    RSLO_CHECK_LAUNCH("rslo_upcat_split");
    return 0;
}

extern "C" int rslo_upcat_backward(const float* dcat, int B, int H, int W, int C, int up, int ld, int choff, float* dz,
                                   int accumulate, rslo_stream_t stream)
{
    if (C % 4 || ld % 4 || choff % 4 || up < 1) return bad("rslo_upcat_backward: bad shape");
    const size_t total = (size_t)B * H * W * (C / 4);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    RSLO_COUNT();
    launch_pdl(k_upcat_bwd, blocks, 256, 0, (cudaStream_t)stream, (const float4*)dcat, B, H, W, C / 4, up, ld / 4, choff / 4, (float4*)dz,
                                                         accumulate);
    RSLO_CHECK_LAUNCH("rslo_upcat_backward");
    return 0;
}

extern "C" int rslo_bias_grad(const float* g, int N, int ld, int C, float* out, rslo_stream_t stream)
{
    if (C < 1 || C > 32) return bad("rslo_bias_grad: C must be in 1..32");
    int blocks = cdiv(N, 8 * 16);
    if (blocks > 148) blocks = 148;
    if (blocks < 1) blocks = 1;
    RSLO_COUNT();
    launch_pdl(k_bias_grad, blocks, 256, 0, (cudaStream_t)stream, g, N, ld, C, out);
    RSLO_CHECK_LAUNCH("rslo_bias_grad");
    return 0;
}

extern "C" int rslo_conv2d_multi_prepare(const rslo_conv_prep_t* table_dev, int n, rslo_stream_t stream)
{
    if (n <= 0) return 0;
    RSLO_COUNT();
    launch_pdl(k_multi_prepare, dim3(24, n), 256, 0, (cudaStream_t)stream, table_dev);
    RSLO_CHECK_LAUNCH("rslo_conv2d_multi_prepare");
    return 0;
}

extern "C" int rslo_conv2d_multi_wgrad_finish(const rslo_wgrad_finish_t* table_dev, int n, rslo_stream_t stream)
{
    if (n <= 0) return 0;
    RSLO_COUNT();
    launch_pdl(k_multi_wgrad_finish, dim3(24, n), 256, 0, (cudaStream_t)stream, table_dev);
    RSLO_CHECK_LAUNCH("rslo_conv2d_multi_wgrad_finish");
    return 0;
}
