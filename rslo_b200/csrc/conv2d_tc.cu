// Dense 2-D convolution of the odometry head on tcgen05 tensor cores, operands staged by TMA (sm_100a).
//
// Replaces cuDNN's FP32 convolutions under `rslo/models/odom_pred_base.py:155-276`,
// `rslo/layers/MaskConv.py:53-63`, `rslo/models/custom_resnet_spc.py:224-298` (3x3 / 1x1, stride 1 / 2).
//
// Activations are NHWC and stored as a *split pair* [2][B][H][W][C]: plane 0 = hi = RN_tf32(x), plane 1 =
// lo = RN_tf32(x - hi) (the tensor core would otherwise TRUNCATE lo's 13 bits to 11: biased, 2x the error).  Weights are prepared once per step as [2][taps][N][Kd] (K-major rows).  The
// convolution is the implicit GEMM
//        out[pixel, n] = sum_tap sum_c A[pixel + offset(tap), c] * Wt[tap][n][c]
// with M = 128 output pixels per CTA (a Wt x Ht patch of one image = one TMA box per tap, zero-filled
// outside the image, which IS the padding), N = NT output channels, K = 32 channels per pipeline stage.
// Split-TF32: lo*hi + hi*lo + hi*hi per stage (FP32-level products, see spconv_tc.cu).
//
//   warp 0 (one lane): TMA producer  - 4 tile loads per stage (A hi/lo: 5-D boxes, W hi/lo: 3-D boxes)
//   warp 1 (one lane): tcgen05.mma issuer (kind::tf32, M128 x N NT x K8), TMEM owner
//   warps 2..5       : drain the double-buffered TMEM accumulator every `group` stages into FP32 registers
//                      (the tensor core accumulates with truncation; long sums are carried in registers),
//                      then bias / split-K combine / store.
//
// One kernel covers forward (stride 1 and 2), data gradient (stride 1; stride 2 as four output-parity
// classes) through a per-launch tap table: each tap = (channel offset, dw, row parity, dh, weight slice)
// in a 5-D view {P*C, W/P, P, H/P, 2B} of the activation (P = 2 turns a stride-2 access into unit-stride
// boxes: even/odd columns are the two halves of a 2C-wide "channel" axis, even/odd rows a size-2 axis).
//
// The weight gradient (k_conv2d_wgrad_tc) contracts over pixels: both operands are MN-major
// (SWIZZLE_128B with 32-byte atoms - the TMA mode of the same name produces exactly that layout).
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>

#include "tma_common.cuh"

namespace rslo {
namespace {
using namespace tc;

constexpr int CV_ROWS = 128;
#ifdef TC_TRACE          // per-CTA / per-stage time stamps (scripts/cv_trace.py); never defined in the shipped build
__device__ unsigned long long g_cv_trace[8 * 2048];
__device__ unsigned long long g_cv_steps[6 * 64];
#define CV_STAMP(slot)                                                                                       \
    do {                                                                                                     \
        const int cta_ = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);                      \
        if (cta_ < 2048) {                                                                                   \
            unsigned long long t_;                                                                           \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                          \
            g_cv_trace[8 * cta_ + (slot)] = t_;                                                              \
        }                                                                                                    \
    } while (0)
#define CV_STEP_STAMP(role, st)                                                                              \
    do {                                                                                                     \
        if (blockIdx.x == 5 && blockIdx.y == 0 && blockIdx.z == 0 && (st) < 64) {                            \
            unsigned long long t_;                                                                           \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                          \
            g_cv_steps[(role) * 64 + (st)] = t_;                                                             \
        }                                                                                                    \
    } while (0)
#else
#define CV_STAMP(slot) do { } while (0)
#define CV_STEP_STAMP(role, st) do { } while (0)
#endif

constexpr int CV_THREADS = 192;
constexpr int CV_KS = 32;                       // channels per stage (one 128-byte swizzle row)
constexpr int CV_A_BYTES = CV_ROWS * CV_KS * 4;  // one plane of the pixel tile: 16 KB
constexpr int CV_MAX_TAPS = 9;

struct ConvTap {
    int coff, dw, p, dh, wslice;
};

struct ConvParams {
    ConvTap taps[CV_MAX_TAPS];
    int ntaps, kchunks, nB, nslices;
    int tiles_w, tiles_h, wt_log2, Wt, Ht;
    int gridW, gridH;
    int OH, OW, osh, ooh, osw, oow, ldo;
    int group;
    int relu;
    int accumulate;                 // epilogue adds to the destination instead of overwriting it
    int stats_c, imgs_per_group;    // BN-statistics epilogue: channels of the stats table, images per statistics group
    int hws[3][3];                  // halo variant: weight slice of tap (dh + 1, dw + 1)
};

// HALO variant (3x3, stride 1): a stage holds, for one column shift dw and one 32-channel chunk, the 8-wide x
// (16 + 2)-high pixel patch ONCE (18 KB per plane) and the three weight slices of that column of taps; the three
// row shifts dh are the same patch read from a start address 8 pixel rows (= one 1024-byte swizzle atom) further
// down, so the activation is fetched 3 times per chunk instead of 9.
constexpr int CV_HALO_W = 8, CV_HALO_H = 16;
constexpr int CV_AH_BYTES = (CV_HALO_H + 2) * CV_HALO_W * CV_KS * 4;     // 18432

template <int NT, bool HALO>
struct CvCfg {
    static constexpr int B_BYTES = NT * CV_KS * 4;
    static constexpr int STAGE_BYTES = HALO ? 2 * CV_AH_BYTES + 6 * B_BYTES : 2 * CV_A_BYTES + 2 * B_BYTES;
    static constexpr int STAGES = HALO ? (NT >= 64 ? 2 : 3) : (NT >= 128 ? 3 : 4);
    static constexpr int TOTAL = STAGES * STAGE_BYTES + 256 + 1024;
    static constexpr int TCOLS = 2 * NT < 32 ? 32 : 2 * NT;
};

__device__ __forceinline__ void drain_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <int NT, bool HALO>
__global__ void __launch_bounds__(CV_THREADS, 1)
k_conv2d_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
            const __grid_constant__ ConvParams P, const float* __restrict__ bias, float* __restrict__ out,
            float* __restrict__ scratch, int* __restrict__ tile_counter, double* __restrict__ stats)
{
    pdl_launch_dependents();
    using S = CvCfg<NT, HALO>;
    constexpr int STAGES = S::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = (uint64_t*)(smem + STAGES * S::STAGE_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + STAGES;
    uint64_t* tfull_bar = bars + 2 * STAGES;
    uint64_t* tempty_bar = bars + 2 * STAGES + 2;
    uint32_t* s_tmem = (uint32_t*)(bars + 2 * STAGES + 4);
    int* s_last = (int*)(s_tmem + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int split = gridDim.y, sidx = blockIdx.y;
    const int n0 = blockIdx.z * NT;
    const int per_img = P.tiles_h * P.tiles_w;
    const int b = blockIdx.x / per_img;
    const int rem = blockIdx.x - b * per_img;
    const int th = rem / P.tiles_w, tw = rem - th * P.tiles_w;
    const int h0 = th * P.Ht, w0 = tw * P.Wt;
    // work items of this CTA: taps x chunks (one stage each), or for HALO (column shift, chunk) pairs = 3 taps each
    const int ntaps_mine = HALO ? 0 : (P.ntaps - sidx + split - 1) / split;
    const int nsteps = HALO ? (3 * P.kchunks - sidx + split - 1) / split : ntaps_mine * P.kchunks;
    const int group = HALO ? 1 : P.group;
    const int ngroups = (nsteps + group - 1) / group;
    if (tid == 0) {
        CV_STAMP(0);
#ifdef TC_TRACE
        g_cv_trace[8 * (blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)) + 6] = nsteps;
#endif
    }

    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(full_bar + i, 1);
            mbar_init(empty_bar + i, 1);
        }
        mbar_init(tfull_bar + 0, 1);
        mbar_init(tfull_bar + 1, 1);
        mbar_init(tempty_bar + 0, 128);
        mbar_init(tempty_bar + 1, 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        tma::prefetch_desc(&tmA);
        tma::prefetch_desc(&tmW);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(S::TCOLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *s_tmem;
    pdl_wait();          // barriers, TMEM and descriptors are set up; global memory is touched from here on
    if (tid == 0) CV_STAMP(1);

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            const uint32_t base = smem_u32(smem);
            int st = 0;
            if constexpr (HALO) {
                for (; st < nsteps; ++st) {
                    const int item = sidx + st * split, dwi = item % 3, kc = item / 3;
                    const int s = st % STAGES;
                    mbar_wait(empty_bar + s, ((st / STAGES) & 1) ^ 1);
                    CV_STEP_STAMP(0, st);                          // stage free
                    const uint32_t a = base + s * S::STAGE_BYTES;
                    tma::mbar_arrive_expect_tx(full_bar + s, S::STAGE_BYTES);
                    tma::load_5d(a, &tmA, kc * CV_KS, w0 + dwi - 1, 0, h0 - 1, b, full_bar + s);
                    tma::load_5d(a + CV_AH_BYTES, &tmA, kc * CV_KS, w0 + dwi - 1, 0, h0 - 1, b + P.nB, full_bar + s);
                    const uint32_t bb = a + 2 * CV_AH_BYTES;
#pragma unroll
                    for (int dhi = 0; dhi < 3; ++dhi) {
                        tma::load_3d(bb + dhi * S::B_BYTES, &tmW, kc * CV_KS, n0, P.hws[dhi][dwi], full_bar + s);
                        tma::load_3d(bb + (3 + dhi) * S::B_BYTES, &tmW, kc * CV_KS, n0, P.hws[dhi][dwi] + P.nslices, full_bar + s);
                    }
                }
            }
            for (int i = 0; i < ntaps_mine; ++i) {
                const ConvTap tp = P.taps[sidx + i * split];
                for (int kc = 0; kc < P.kchunks; ++kc, ++st) {
                    const int s = st % STAGES;
                    mbar_wait(empty_bar + s, ((st / STAGES) & 1) ^ 1);
                    const uint32_t a = base + s * S::STAGE_BYTES;
                    tma::mbar_arrive_expect_tx(full_bar + s, S::STAGE_BYTES);
                    tma::load_5d(a, &tmA, tp.coff + kc * CV_KS, w0 + tp.dw, tp.p, h0 + tp.dh, b, full_bar + s);
                    tma::load_5d(a + CV_A_BYTES, &tmA, tp.coff + kc * CV_KS, w0 + tp.dw, tp.p, h0 + tp.dh, b + P.nB,
                                 full_bar + s);
                    tma::load_3d(a + 2 * CV_A_BYTES, &tmW, kc * CV_KS, n0, tp.wslice, full_bar + s);
                    tma::load_3d(a + 2 * CV_A_BYTES + S::B_BYTES, &tmW, kc * CV_KS, n0, tp.wslice + P.nslices, full_bar + s);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_tf32(NT);
            uint32_t acc = 0;
            if constexpr (HALO) {
                for (int st = 0; st < nsteps; ++st) {
                    const int buf = st & 1, s = st % STAGES;
                    mbar_wait(tempty_bar + buf, ((st >> 1) & 1) ^ 1);
                    CV_STEP_STAMP(1, st);                          // accumulator buffer free
                    mbar_wait(full_bar + s, (st / STAGES) & 1);
                    CV_STEP_STAMP(2, st);                          // operands landed
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t d = tmem_base + buf * NT;
                    const uint32_t sb = smem_u32(smem) + s * S::STAGE_BYTES;
                    acc = 0;
#pragma unroll
                    for (int dhi = 0; dhi < 3; ++dhi) {
                        const uint32_t a_hi = sb + dhi * (CV_HALO_W * CV_KS * 4), a_lo = a_hi + CV_AH_BYTES;
                        const uint32_t b_hi = sb + 2 * CV_AH_BYTES + dhi * S::B_BYTES, b_lo = b_hi + 3 * S::B_BYTES;
#pragma unroll
                        for (int part = 0; part < 3; ++part) {        // lo*hi, hi*lo, hi*hi
                            const uint32_t a = part == 0 ? a_lo : a_hi;
                            const uint32_t bb = part == 1 ? b_lo : b_hi;
#pragma unroll
                            for (int kk = 0; kk < CV_KS / 8; ++kk) {
                                umma_tf32(d, umma_desc_k_sw128(a + kk * 32), umma_desc_k_sw128(bb + kk * 32), idesc, acc);
                                acc = 1;
                            }
                        }
                    }
                    umma_commit(empty_bar + s);
                    umma_commit(tfull_bar + buf);
                    CV_STEP_STAMP(3, st);                          // MMAs issued
                }
            }
            for (int st = 0; st < (HALO ? 0 : nsteps); ++st) {
                const int grp = st / P.group, buf = grp & 1;
                if (st - grp * P.group == 0) {
                    mbar_wait(tempty_bar + buf, ((grp >> 1) & 1) ^ 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    acc = 0;
                }
                const int s = st % STAGES;
                mbar_wait(full_bar + s, (st / STAGES) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d = tmem_base + buf * NT;
                const uint32_t a_hi = smem_u32(smem) + s * S::STAGE_BYTES;
                const uint32_t a_lo = a_hi + CV_A_BYTES;
                const uint32_t b_hi = a_hi + 2 * CV_A_BYTES;
                const uint32_t b_lo = b_hi + S::B_BYTES;
#pragma unroll
                for (int part = 0; part < 3; ++part) {            // lo*hi, hi*lo, hi*hi
                    const uint32_t a = part == 0 ? a_lo : a_hi;
                    const uint32_t bb = part == 1 ? b_lo : b_hi;
#pragma unroll
                    for (int kk = 0; kk < CV_KS / 8; ++kk) {
                        umma_tf32(d, umma_desc_k_sw128(a + kk * 32), umma_desc_k_sw128(bb + kk * 32), idesc, acc);
                        acc = 1;
                    }
                }
                umma_commit(empty_bar + s);
                if (st + 1 == nsteps || (st + 1) % P.group == 0) umma_commit(tfull_bar + buf);
            }
        }
        __syncwarp();
    } else {
        // ================= drain + epilogue (one output pixel per thread) =================
        const int q = warp & 3;
        const int r = q * 32 + lane;
        float acc[NT];
#pragma unroll
        for (int i = 0; i < NT; ++i) acc[i] = 0.f;
        for (int g = 0; g < ngroups; ++g) {
            const int buf = g & 1;
            mbar_wait(tfull_bar + buf, (g >> 1) & 1);
            if (warp == 2 && lane == 0) CV_STEP_STAMP(4, g);       // MMAs of the stage retired
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * NT;
#pragma unroll
            for (int cb = 0; cb < NT; cb += 16) {
                float v[16];
                tmem_ld16(taddr + cb, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[cb + i] += v[i];
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(tempty_bar + buf);
            if (warp == 2 && lane == 0) CV_STEP_STAMP(5, g);       // drained
        }
        if (warp == 2 && lane == 0) CV_STAMP(4);
        bool finish = true;
        if (split > 1) {
            const int wtile = blockIdx.x * gridDim.z + blockIdx.z;
            float4* mine = reinterpret_cast<float4*>(scratch + ((size_t)(wtile * split + sidx) * CV_ROWS + r) * NT);
#pragma unroll
            for (int i = 0; i < NT; i += 4) mine[i / 4] = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
            __threadfence();
            drain_sync();
            if (warp == 2 && lane == 0) *s_last = atomicAdd(tile_counter + wtile, 1) == split - 1;
            drain_sync();
            finish = *s_last != 0;
            if (finish) {
                __threadfence();
#pragma unroll
                for (int i = 0; i < NT; ++i) acc[i] = 0.f;
                for (int sp = 0; sp < split; ++sp) {               // fixed order: deterministic sum
                    const float4* p = reinterpret_cast<const float4*>(scratch + ((size_t)(wtile * split + sp) * CV_ROWS + r) * NT);
#pragma unroll
                    for (int i = 0; i < NT; i += 4) {
                        const float4 t = __ldcg(p + i / 4);
                        acc[i] += t.x; acc[i + 1] += t.y; acc[i + 2] += t.z; acc[i + 3] += t.w;
                    }
                }
                if (warp == 2 && lane == 0) tile_counter[wtile] = 0;      // leave the counters zeroed for the next launch
            }
        }
        const int wl = r & (P.Wt - 1), hl = r >> P.wt_log2;
        const int h = h0 + hl, w = w0 + wl;
        const bool valid = h < P.gridH && w < P.gridW;
        if (finish) {
            if (bias) {
#pragma unroll
                for (int i = 0; i < NT; i += 4) {
                    const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + n0 + i));
                    acc[i] += bv.x; acc[i + 1] += bv.y; acc[i + 2] += bv.z; acc[i + 3] += bv.w;
                }
            }
            if (valid) {
                float4* dst = reinterpret_cast<float4*>(
                    out + (((size_t)b * P.OH + (size_t)(h * P.osh + P.ooh)) * P.OW + (size_t)(w * P.osw + P.oow)) * P.ldo + n0);
#pragma unroll
                for (int i = 0; i < NT; i += 4) {
                    float4 v = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
                    if (P.relu) {
                        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
                    }
                    if (P.accumulate) {
                        const float4 o = dst[i / 4];
                        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
                    }
                    dst[i / 4] = v;
                }
            }
            if (stats != nullptr) {
                // per-channel sum / sum of squares of this tile's valid pixels (BatchNorm batch statistics of the
                // statistics group the image belongs to): transpose through the now idle pipeline stages
                float* tile = reinterpret_cast<float*>(smem);
                constexpr int LD = NT + 1;
#pragma unroll
                for (int i = 0; i < NT; ++i) tile[r * LD + i] = valid ? acc[i] : 0.f;
                drain_sync();
                double* dst = stats + ((size_t)(b / P.imgs_per_group) * P.stats_c + n0) * 2;
                for (int c = r; c < NT; c += CV_ROWS) {
                    float s1 = 0.f, s2 = 0.f;
#pragma unroll 8
                    for (int row = 0; row < CV_ROWS; ++row) {
                        const float v = tile[row * LD + c];
                        s1 += v;
                        s2 = fmaf(v, v, s2);
                    }
                    atomicAdd(dst + 2 * c, (double)s1);
                    atomicAdd(dst + 2 * c + 1, (double)s2);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) CV_STAMP(5);
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(S::TCOLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
// pick the pixel-tile shape (Wt x Ht = npix, powers of two) that covers a gridW x gridH image with the fewest tiles
static void pick_tile(int gridW, int gridH, int npix, int* Wt, int* Ht)
{
    long best = -1;
    for (int wt = npix; wt >= 1; wt >>= 1) {
        const int ht = npix / wt;
        if (wt > 256 || ht > 256) continue;
        const long tiles = (long)cdiv(gridW, wt) * cdiv(gridH, ht);
        if (best < 0 || tiles < best) {
            best = tiles;
            *Wt = wt;
            *Ht = ht;
        }
    }
}
static int ilog2(int v)
{
    int l = 0;
    while ((1 << l) < v) ++l;
    return l;
}

// 5-D view {P*C, W/P, P, H/P, 2B} of a split-pair activation [2][B][H][W][C]
static int encode_act(CUtensorMap* tm, const float* base, int B, int H, int W, int C, int P, int box_w, int box_h,
                      CUtensorMapSwizzle sw)
{
    const uint64_t dims[5] = {(uint64_t)P * C, (uint64_t)(W / P), (uint64_t)P, (uint64_t)(H / P), (uint64_t)2 * B};
    const uint64_t strides[4] = {(uint64_t)P * C * 4, (uint64_t)W * C * 4, (uint64_t)P * W * C * 4, (uint64_t)H * W * C * 4};
    const uint32_t box[5] = {32, (uint32_t)box_w, 1, (uint32_t)box_h, 1};
    return tma::encode_f32(tm, base, 5, dims, strides, box, sw);
}

template <int NT, bool HALO>
static int launch_conv(const CUtensorMap& tmA, const CUtensorMap& tmW, const ConvParams& P, int tiles, int split, int ntiles,
                       const float* bias, float* out, float* scratch, int* counter, double* stats, cudaStream_t st)
{
    using S = CvCfg<NT, HALO>;
    static bool configured = false;
    if (!configured) {
        RSLO_CHECK(cudaFuncSetAttribute(k_conv2d_tc<NT, HALO>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
        configured = true;
    }
    RSLO_COUNT();
    launch_pdl(k_conv2d_tc<NT, HALO>, dim3(tiles, split, ntiles), CV_THREADS, S::TOTAL, st, tmA, tmW, P, bias, out, scratch, counter,
                                                                                 stats);
    RSLO_CHECK_LAUNCH("rslo_conv2d_tc");
    return 0;
}

static int pick_nt(int N, int pixel_tiles)
{
    if (N % 128 == 0 && (long)pixel_tiles * (N / 128) >= 96) return 128;
    if (N % 64 == 0) return 64;
    if (N % 32 == 0) return 32;
    return 0;
}
// RSLO_CONV_HALO: 1 = always the halo variant where it applies, 0 = never, unset = measure both once per layer shape
static int conv_halo_mode()
{
    static int v = -2;
    if (v == -2) {
        const char* e = getenv("RSLO_CONV_HALO");
        v = e ? (atoi(e) != 0 ? 1 : 0) : -1;
    }
    return v;
}
static int conv_drain_group()
{
    static int g = 0;
    if (g == 0) {
        const char* e = getenv("RSLO_CONV_DRAIN_GROUP");      // measurement switch
        g = e ? atoi(e) : 1;
        if (g < 1) g = 1;
        if (g > 8) g = 8;
    }
    return g;
}
static int pick_split(int ctas, int ntaps)
{
    if (ntaps >= 9 && ctas * 9 <= 160) return 9;
    if (ntaps >= 3 && ctas * 3 <= 160) return 3;
    return 1;
}

// workspace layout of the split-K combine: [counters: 4096 ints][scratch floats]
constexpr int CV_MAX_COUNTERS = 4096;

// Generic implicit-GEMM launch: A = split-pair activation [2][B][H][W][C] viewed with parity factor Pf,
// weight image [2][nslices][N][C], output pixel grid gridH x gridW per image (tile coordinates),
// taps in tile coordinates, output address mapping (OH, OW, osh, ooh, osw, oow, ldo).
static bool halo_eligible(int Pf, const ConvTap* taps, int ntaps)
{
    if (Pf != 1 || ntaps != 9) return false;
    int seen = 0;
    for (int i = 0; i < 9; ++i) {
        const ConvTap& t = taps[i];
        if (t.coff != 0 || t.p != 0 || t.dw < -1 || t.dw > 1 || t.dh < -1 || t.dh > 1) return false;
        seen |= 1 << ((t.dh + 1) * 3 + t.dw + 1);
    }
    return seen == 0x1ff;
}

static int run_conv_impl(bool halo, const float* a_split, int B, int H, int W, int C, int Pf, const float* wimg, int nslices,
                         int N, int gridH, int gridW, const ConvTap* taps, int ntaps, float* out, int OH, int OW, int osh,
                         int ooh, int osw, int oow, int ldo, const float* bias, int relu, void* ws, size_t ws_bytes,
                         cudaStream_t st, double* stats, int imgs_per_group, int accumulate)
{
    if (C % 32 != 0 || ntaps < 1 || ntaps > CV_MAX_TAPS || (H % Pf) || (W % Pf)) {
        set_last_error("rslo_conv2d_tc: unsupported shape", cudaErrorInvalidValue);
        return (int)cudaErrorInvalidValue;
    }
    ConvParams P;
    memset(&P, 0, sizeof P);
    // 3x3 / stride 1 (forward, and the data gradient of such a layer): halo variant, see CvCfg
    if (halo) {
        for (int i = 0; i < 9; ++i) P.hws[taps[i].dh + 1][taps[i].dw + 1] = taps[i].wslice;
        P.Wt = CV_HALO_W;
        P.Ht = CV_HALO_H;
    } else {
        pick_tile(gridW, gridH, CV_ROWS, &P.Wt, &P.Ht);
    }
    P.wt_log2 = ilog2(P.Wt);
    P.tiles_w = cdiv(gridW, P.Wt);
    P.tiles_h = cdiv(gridH, P.Ht);
    const int tiles = B * P.tiles_w * P.tiles_h;
    int NT = pick_nt(N, tiles);
    if (NT == 0) {
        set_last_error("rslo_conv2d_tc: output channels must be a multiple of 32", cudaErrorInvalidValue);
        return (int)cudaErrorInvalidValue;
    }
    if (halo && NT > 64) NT = 64;                        // three weight slices per stage: 64 output channels fit
    const int ntiles = N / NT;
    int split = pick_split(tiles * ntiles, ntaps);
    if (halo) {                                          // work items = (column shift, chunk) pairs
        const int items = 3 * (C / CV_KS);
        split = 1;
        for (int d = items; d > 1; --d)
            if (items % d == 0 && (long)tiles * ntiles * d <= 160) {
                split = d;
                break;
            }
    }
    for (int i = 0; i < ntaps; ++i) P.taps[i] = taps[i];
    P.ntaps = ntaps;
    P.kchunks = C / CV_KS;
    P.nB = B;
    P.nslices = nslices;
    P.gridW = gridW;
    P.gridH = gridH;
    P.OH = OH; P.OW = OW; P.osh = osh; P.ooh = ooh; P.osw = osw; P.oow = oow; P.ldo = ldo;
    // stages accumulated in TMEM between two drains.  The tensor core adds into its FP32 accumulator with
    // truncation: ~1e-8 relative bias per add, always towards zero, and with batch statistics switched off
    // (eval / frozen BatchNorm) that bias survives all ~35 layers.  One stage = 12 adds keeps a layer at ~1e-7.
    P.group = conv_drain_group() < P.kchunks ? conv_drain_group() : P.kchunks;
    P.relu = relu;
    P.accumulate = accumulate;
    P.stats_c = N;
    P.imgs_per_group = imgs_per_group > 0 ? imgs_per_group : 1;

    float* scratch = nullptr;
    int* counter = nullptr;
    if (split > 1) {
        const size_t need = (size_t)CV_MAX_COUNTERS * 4 + (size_t)tiles * ntiles * split * CV_ROWS * NT * 4;
        if (tiles * ntiles > CV_MAX_COUNTERS || ws == nullptr || ws_bytes < need) split = 1;
        else {
            counter = (int*)ws;                       // zeroed once by the caller; kernels re-zero what they use
            scratch = (float*)((char*)ws + (size_t)CV_MAX_COUNTERS * 4);
        }
    }
    CUtensorMap tmA, tmW;
    int rc = encode_act(&tmA, a_split, B, H, W, C, Pf, P.Wt, halo ? P.Ht + 2 : P.Ht, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    {
        const uint64_t dims[3] = {(uint64_t)C, (uint64_t)N, (uint64_t)2 * nslices};
        const uint64_t strides[2] = {(uint64_t)C * 4, (uint64_t)N * C * 4};
        const uint32_t box[3] = {32, (uint32_t)NT, 1};
        rc = tma::encode_f32(&tmW, wimg, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    if (halo) {
        if (NT == 64) return launch_conv<64, true>(tmA, tmW, P, tiles, split, ntiles, bias, out, scratch, counter, stats, st);
        return launch_conv<32, true>(tmA, tmW, P, tiles, split, ntiles, bias, out, scratch, counter, stats, st);
    }
    if (NT == 128) return launch_conv<128, false>(tmA, tmW, P, tiles, split, ntiles, bias, out, scratch, counter, stats, st);
    if (NT == 64) return launch_conv<64, false>(tmA, tmW, P, tiles, split, ntiles, bias, out, scratch, counter, stats, st);
    return launch_conv<32, false>(tmA, tmW, P, tiles, split, ntiles, bias, out, scratch, counter, stats, st);
}

// Which variant runs a 3x3 / stride-1 layer: fewer activation bytes (halo) usually wins, but wide data gradients
// (N >= 256: many 64-channel tiles re-reading the patch) can favour the plain tiling with its deeper pipeline.  Each layer
// shape is measured once (both variants, CUDA events on the launching stream, first non-accumulating call outside graph
// capture) and the choice cached for the life of the process.
struct TuneKey {
    int v[8];
    bool operator<(const TuneKey& o) const { return memcmp(v, o.v, sizeof v) < 0; }
};
static std::map<TuneKey, int> g_conv_choice;
static std::mutex g_conv_mutex;

static int run_conv(const float* a_split, int B, int H, int W, int C, int Pf, const float* wimg, int nslices, int N,
                    int gridH, int gridW, const ConvTap* taps, int ntaps, float* out, int OH, int OW, int osh, int ooh,
                    int osw, int oow, int ldo, const float* bias, int relu, void* ws, size_t ws_bytes, cudaStream_t st,
                    double* stats = nullptr, int imgs_per_group = 1, int accumulate = 0)
{
#define RUN(h, o_stats, o_acc)                                                                                              \
    run_conv_impl(h, a_split, B, H, W, C, Pf, wimg, nslices, N, gridH, gridW, taps, ntaps, out, OH, OW, osh, ooh, osw, oow,  \
                  ldo, bias, relu, ws, ws_bytes, st, o_stats, imgs_per_group, o_acc)
    const int mode = conv_halo_mode();
    if (mode == 0 || C % 32 != 0 || !halo_eligible(Pf, taps, ntaps)) return RUN(false, stats, accumulate);
    if (mode == 1) return RUN(true, stats, accumulate);
    const TuneKey key = {{B, H, W, C, N, gridH, gridW, ws != nullptr}};
    int choice = -1;
    {
        std::lock_guard<std::mutex> lock(g_conv_mutex);
        auto it = g_conv_choice.find(key);
        if (it != g_conv_choice.end()) choice = it->second;
    }
    if (choice < 0) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(st, &cap);
        if (cap != cudaStreamCaptureStatusNone || accumulate) return RUN(N <= 128, stats, accumulate);   // not measurable now
        cudaEvent_t ev[3];
        for (auto& e : ev) RSLO_CHECK(cudaEventCreate(&e));
        float ms[2] = {0.f, 0.f};
        for (int h = 0; h < 2; ++h) {
            int rc = RUN(h == 1, nullptr, 0);                              // warm (tensor maps, instruction cache)
            if (rc) return rc;
            RSLO_CHECK(cudaEventRecord(ev[h], st));
            for (int i = 0; i < 3; ++i)
                if ((rc = RUN(h == 1, nullptr, 0))) return rc;
            RSLO_CHECK(cudaEventRecord(ev[h + 1], st));
            if (h == 0) {                                                  // ev[1] is reused as the second start
                RSLO_CHECK(cudaEventSynchronize(ev[1]));
                RSLO_CHECK(cudaEventElapsedTime(&ms[0], ev[0], ev[1]));
            }
        }
        RSLO_CHECK(cudaEventSynchronize(ev[2]));
        RSLO_CHECK(cudaEventElapsedTime(&ms[1], ev[1], ev[2]));
        for (auto& e : ev) cudaEventDestroy(e);
        choice = ms[1] < 0.97f * ms[0] ? 1 : 0;
        std::lock_guard<std::mutex> lock(g_conv_mutex);
        g_conv_choice[key] = choice;
    }
    return RUN(choice == 1, stats, accumulate);
#undef RUN
}

// forward taps of a ks x ks / stride s / pad (ks/2) convolution read through the parity view (Pf = s)
static int forward_taps(int ks, int s, int C, ConvTap* taps)
{
    const int pad = ks / 2;
    int n = 0;
    for (int ky = 0; ky < ks; ++ky)
        for (int kx = 0; kx < ks; ++kx) {
            const int dy = ky - pad, dx = kx - pad;
            ConvTap t;
            if (s == 1) {
                t.coff = 0; t.dw = dx; t.p = 0; t.dh = dy;
            } else {
                const int py = dy & 1, px = dx & 1;
                t.p = py; t.dh = (dy - py) / 2;
                t.coff = px * C; t.dw = (dx - px) / 2;
            }
            t.wslice = ky * ks + kx;
            taps[n++] = t;
        }
    return n;
}

// ---- elementwise helpers ---------------------------------------------------------------------------
__global__ void k_split_planes(const float4* __restrict__ x, size_t n4, float4* __restrict__ hi, float4* __restrict__ lo)
{
    pdl_launch_dependents();
    pdl_wait();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(x + i);
        const float4 h = make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
        hi[i] = h;
        lo[i] = make_float4(tf32_rn(v.x - h.x), tf32_rn(v.y - h.y), tf32_rn(v.z - h.z), tf32_rn(v.w - h.w));
    }
}

// OIHW weight -> split image [2][taps][N][Kd]:  mode 0 (forward): N = Cout, Kd = Cin, img[t][co][ci];
// mode 1 (data gradient): N = Cin, Kd = Cout, img[t][ci][co]
// (output channels zero-padded to CoutP, a multiple of 32, for the narrow 7- / 1-channel heads)
__global__ void k_conv2d_wprep(const float* __restrict__ w, int Cout, int CoutP, int Cin, int taps, int mode,
                               float* __restrict__ img)
{
    pdl_launch_dependents();
    pdl_wait();
    const int total = taps * CoutP * Cin;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int N = mode ? Cin : CoutP, Kd = mode ? CoutP : Cin;
    const int kk = i % Kd, n = (i / Kd) % N, t = i / (Kd * N);
    const int co = mode ? kk : n, ci = mode ? n : kk;
    const float v = co < Cout ? __ldg(w + ((size_t)co * Cin + ci) * taps + t) : 0.f;
    const float h = tf32_rn(v);
    img[i] = h;
    img[(size_t)total + i] = tf32_rn(v - h);
}

// ---------------------------------------------------------------------------------------------------
// weight gradient: dW[tap][ci][co] += sum_pixels X[pixel + offset(tap)][ci] * G[pixel][co]
// ---------------------------------------------------------------------------------------------------
constexpr int WG_PIX = 32;                          // pixels (K) per stage
constexpr int WG_CHUNK_BYTES = WG_PIX * 128;        // one [32 px x 32 ch] box: 4 KB
constexpr int WG_A_BYTES = 2 * 4 * WG_CHUNK_BYTES;  // {hi, lo} x 4 chunks = 32 KB
constexpr int WG_A_STAGES = 4;
constexpr int WG_G_STAGES = 2;
constexpr int WG_THREADS = 192;

struct WgradParams {
    ConvTap taps[CV_MAX_TAPS];
    int ntaps, cpt;                 // chunks (of 32 input channels) per tap
    int nchunks;                    // ntaps * cpt
    int nB;                         // images
    int tiles_w, tiles_h, Wt, Ht;   // 32-pixel tiles over the OUTPUT grid
    int steps_total, steps_per_cta;
    int groups_per_cta, ngroups;
    int Cin, Cout;
};

template <int NT>
struct WgCfg {
    static constexpr int G_BYTES = 2 * (NT / 32) * WG_CHUNK_BYTES;
    static constexpr int TOTAL = WG_A_STAGES * WG_A_BYTES + WG_G_STAGES * G_BYTES + 256 + 1024;
    static constexpr int MAXG = 512 / NT;
};

__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t umma_idesc_tf32_mn(int n)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(128 >> 4) << 24);
}

// grid = (pixel splits, group sets, N tiles)
template <int NT>
__global__ void __launch_bounds__(WG_THREADS, 1)
k_conv2d_wgrad_tc(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG,
                  const __grid_constant__ WgradParams P, float* __restrict__ dW)
{
    pdl_launch_dependents();
    using S = WgCfg<NT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sG = smem + WG_A_STAGES * WG_A_BYTES;
    uint64_t* bars = (uint64_t*)(sG + WG_G_STAGES * S::G_BYTES);
    uint64_t* a_full = bars;                       // [WG_A_STAGES]
    uint64_t* a_empty = bars + WG_A_STAGES;        // [WG_A_STAGES]
    uint64_t* g_full = bars + 2 * WG_A_STAGES;     // [WG_G_STAGES]
    uint64_t* g_empty = g_full + WG_G_STAGES;      // [WG_G_STAGES]
    uint64_t* acc_bar = g_empty + WG_G_STAGES;
    uint32_t* s_tmem = (uint32_t*)(acc_bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int step0 = blockIdx.x * P.steps_per_cta;
    const int step1 = min(P.steps_total, step0 + P.steps_per_cta);
    const int g0 = blockIdx.y * P.groups_per_cta;
    const int ng = min(P.groups_per_cta, P.ngroups - g0);
    const int n0 = blockIdx.z * NT;
    if (step0 >= step1 || ng <= 0) return;
    constexpr int TCOLS = 512;

    if (tid == 0) {
        for (int i = 0; i < WG_A_STAGES; ++i) {
            mbar_init(a_full + i, 1);
            mbar_init(a_empty + i, 1);
        }
        for (int i = 0; i < WG_G_STAGES; ++i) {
            mbar_init(g_full + i, 1);
            mbar_init(g_empty + i, 1);
        }
        mbar_init(acc_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        tma::prefetch_desc(&tmX);
        tma::prefetch_desc(&tmG);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(TCOLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *s_tmem;
    pdl_wait();          // barriers, TMEM and descriptors are set up; global memory is touched from here on
    const int per_img = P.tiles_h * P.tiles_w;

    if (warp == 0) {
        if (elect_one()) {
            const uint32_t sA_u = smem_u32(sA), sG_u = smem_u32(sG);
            int item = 0;
            for (int stp = step0; stp < step1; ++stp) {
                const int b = stp / per_img;
                const int rem = stp - b * per_img;
                const int th = rem / P.tiles_w, tw = rem - th * P.tiles_w;
                const int h0 = th * P.Ht, w0 = tw * P.Wt;
                {
                    const int it = stp - step0, gs = it % WG_G_STAGES;
                    mbar_wait(g_empty + gs, ((it / WG_G_STAGES) & 1) ^ 1);
                    const uint32_t gb = sG_u + gs * S::G_BYTES;
                    tma::mbar_arrive_expect_tx(g_full + gs, S::G_BYTES);
#pragma unroll
                    for (int c = 0; c < NT / 32; ++c) {
                        tma::load_5d(gb + c * WG_CHUNK_BYTES, &tmG, n0 + c * 32, w0, 0, h0, b, g_full + gs);
                        tma::load_5d(gb + (NT / 32 + c) * WG_CHUNK_BYTES, &tmG, n0 + c * 32, w0, 0, h0, b + P.nB, g_full + gs);
                    }
                }
                for (int gi = 0; gi < ng; ++gi, ++item) {
                    const int as = item % WG_A_STAGES;
                    mbar_wait(a_empty + as, ((item / WG_A_STAGES) & 1) ^ 1);
                    const uint32_t ab = sA_u + as * WG_A_BYTES;
                    const int j0 = (g0 + gi) * 4;
                    const int nch = min(4, P.nchunks - j0);
                    tma::mbar_arrive_expect_tx(a_full + as, (uint32_t)nch * 2 * WG_CHUNK_BYTES);
                    for (int c = 0; c < nch; ++c) {
                        const int j = j0 + c;
                        const int t = j / P.cpt, cc = j - t * P.cpt;
                        const ConvTap tp = P.taps[t];
                        tma::load_5d(ab + c * WG_CHUNK_BYTES, &tmX, tp.coff + cc * 32, w0 + tp.dw, tp.p, h0 + tp.dh, b, a_full + as);
                        tma::load_5d(ab + (4 + c) * WG_CHUNK_BYTES, &tmX, tp.coff + cc * 32, w0 + tp.dw, tp.p, h0 + tp.dh, b + P.nB,
                                     a_full + as);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_tf32_mn(NT);
            constexpr uint32_t LBO = WG_CHUNK_BYTES;
            int item = 0;
            for (int stp = step0; stp < step1; ++stp) {
                const int it = stp - step0, gs = it % WG_G_STAGES;
                mbar_wait(g_full + gs, (it / WG_G_STAGES) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t gb = smem_u32(sG) + gs * S::G_BYTES;
                for (int gi = 0; gi < ng; ++gi, ++item) {
                    const int as = item % WG_A_STAGES;
                    mbar_wait(a_full + as, (item / WG_A_STAGES) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t ab = smem_u32(sA) + as * WG_A_BYTES;
                    const uint32_t d = tmem_base + gi * NT;
#pragma unroll
                    for (int part = 0; part < 3; ++part) {
                        const uint32_t a = ab + (part == 0 ? 4 * WG_CHUNK_BYTES : 0);
                        const uint32_t bb = gb + (part == 1 ? (NT / 32) * WG_CHUNK_BYTES : 0);
#pragma unroll
                        for (int kg = 0; kg < WG_PIX / 8; ++kg) {
                            umma_tf32(d, umma_desc_mn_sw128(a + kg * 1024, LBO, 512), umma_desc_mn_sw128(bb + kg * 1024, LBO, 512),
                                      idesc, (it > 0 || part > 0 || kg > 0) ? 1u : 0u);
                        }
                    }
                    umma_commit(a_empty + as);
                }
                umma_commit(g_empty + gs);
            }
            umma_commit(acc_bar);
        }
        __syncwarp();
    } else {
        const int q = warp & 3;
        const int m = q * 32 + lane;
        mbar_wait(acc_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int gi = 0; gi < ng; ++gi) {
            const int j = (g0 + gi) * 4 + (m >> 5);
            const bool ok = j < P.nchunks;
            const int t = ok ? j / P.cpt : 0;
            const int ci = ok ? (j - t * P.cpt) * 32 + (m & 31) : 0;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + gi * NT;
#pragma unroll
            for (int cb = 0; cb < NT; cb += 16) {
                float v[16];
                tmem_ld16(taddr + cb, v);
                if (ok) {
                    float* dst = dW + ((size_t)t * P.Cin + ci) * P.Cout + n0 + cb;
#pragma unroll
                    for (int i = 0; i < 16; i += 4)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "f"(v[i]), "f"(v[i + 1]),
                                     "f"(v[i + 2]), "f"(v[i + 3])
                                     : "memory");
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TCOLS) : "memory");
    }
}

template <int NT>
static int launch_wgrad(const CUtensorMap& tmX, const CUtensorMap& tmG, const WgradParams& P, dim3 grid, float* dW,
                        cudaStream_t st)
{
    using S = WgCfg<NT>;
    static bool configured = false;
    if (!configured) {
        RSLO_CHECK(cudaFuncSetAttribute(k_conv2d_wgrad_tc<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
        configured = true;
    }
    RSLO_COUNT();
    launch_pdl(k_conv2d_wgrad_tc<NT>, grid, WG_THREADS, S::TOTAL, st, tmX, tmG, P, dW);
    RSLO_CHECK_LAUNCH("rslo_conv2d_tc_backward_weight");
    return 0;
}

// dW [taps][Cin][Cout] -> OIHW gradient (+= when accumulate)
__global__ void k_wgrad_finish(const float* __restrict__ dW, int Cout, int CoutP, int Cin, int taps, int accumulate,
                               float* __restrict__ gw)
{
    pdl_launch_dependents();
    pdl_wait();
    const int total = taps * Cout * Cin;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;         // OIHW index (real output channels only)
    if (i >= total) return;
    const int t = i % taps, ci = (i / taps) % Cin, co = i / (taps * Cin);
    const float v = __ldg(dW + ((size_t)t * Cin + ci) * CoutP + co);
    gw[i] = accumulate ? gw[i] + v : v;
}

}  // namespace
}  // namespace rslo

using namespace rslo;

#ifdef TC_TRACE
extern "C" int rslo_debug_cv_trace(unsigned long long* ctas, unsigned long long* steps)
{
    cudaMemcpyFromSymbol(ctas, g_cv_trace, sizeof(g_cv_trace));
    return (int)cudaMemcpyFromSymbol(steps, g_cv_steps, sizeof(g_cv_steps));
}
extern "C" int rslo_debug_cv_clear()
{
    static unsigned long long z[8 * 2048];
    cudaMemcpyToSymbol(g_cv_trace, z, sizeof(g_cv_trace));
    return (int)cudaMemcpyToSymbol(g_cv_steps, z, sizeof(g_cv_steps));
}
#endif

extern "C" int rslo_conv2d_tc_supported(int Cin, int Cout, int ksize, int stride)
{
    return Cin % 32 == 0 && Cout % 32 == 0 && Cin >= 32 && Cout >= 32 && (ksize == 1 || ksize == 3) && (stride == 1 || stride == 2);
}

extern "C" int rslo_conv2d_split(const float* x, size_t n, float* split_pair, rslo_stream_t stream)
{
    if (n == 0) return 0;
    if (n % 4 != 0) {
        set_last_error("rslo_conv2d_split: element count must be a multiple of 4", cudaErrorInvalidValue);
        return (int)cudaErrorInvalidValue;
    }
    const size_t n4 = n / 4;
    int blocks = (int)((n4 + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    RSLO_COUNT();
    launch_pdl(k_split_planes, blocks, 256, 0, (cudaStream_t)stream, (const float4*)x, n4, (float4*)split_pair, (float4*)(split_pair + n));
    RSLO_CHECK_LAUNCH("rslo_conv2d_split");
    return 0;
}

extern "C" int rslo_conv2d_tc_prepare(const float* weight_oihw, int Cout, int Cout_padded, int Cin, int ksize, int mode,
                                      float* image, rslo_stream_t stream)
{
    if (Cout_padded < Cout) Cout_padded = Cout;
    const int total = ksize * ksize * Cout_padded * Cin;
    RSLO_COUNT();
    launch_pdl(k_conv2d_wprep, cdiv(total, 256), 256, 0, (cudaStream_t)stream, weight_oihw, Cout, Cout_padded, Cin, ksize * ksize,
                                                                      mode, image);
    RSLO_CHECK_LAUNCH("rslo_conv2d_tc_prepare");
    return 0;
}

extern "C" size_t rslo_conv2d_tc_workspace_bytes(int B, int H, int W, int Cmax)
{
    // split-K scratch is only used by launches with < 160/3 CTAs: bound = 160 CTAs x 128 rows x 128 channels
    (void)B; (void)H; (void)W; (void)Cmax;
    return (size_t)CV_MAX_COUNTERS * 4 + (size_t)160 * CV_ROWS * 128 * 4 + 256;
}

extern "C" int rslo_conv2d_tc_forward(const float* x_split, int B, int H, int W, int Cin, const float* image, int Cout,
                                      int ksize, int stride, const float* bias, int relu, float* y, double* stats,
                                      int imgs_per_group, void* workspace, size_t workspace_bytes, rslo_stream_t stream)
{
    if (!rslo_conv2d_tc_supported(Cin, Cout, ksize, stride)) {
        set_last_error("rslo_conv2d_tc_forward: unsupported shape", cudaErrorInvalidValue);
        return (int)cudaErrorInvalidValue;
    }
    const int pad = ksize / 2;
    const int Ho = (H + 2 * pad - ksize) / stride + 1, Wo = (W + 2 * pad - ksize) / stride + 1;
    ConvTap taps[CV_MAX_TAPS];
    const int nt = forward_taps(ksize, stride, Cin, taps);
    return run_conv(x_split, B, H, W, Cin, stride, image, ksize * ksize, Cout, Ho, Wo, taps, nt, y, Ho, Wo, 1, 0, 1, 0, Cout,
                    bias, relu, workspace, workspace_bytes, (cudaStream_t)stream, stats, imgs_per_group);
}

// dx [B][H][W][Cin] from g_split [2][B][Ho][Wo][Cout]; image = mode-1 prepared weights [2][taps][Cin][Cout]
extern "C" int rslo_conv2d_tc_backward_data(const float* g_split, int B, int H, int W, int Cin, const float* image_t, int Cout,
                                            int ksize, int stride, float* dx, int accumulate, void* workspace,
                                            size_t workspace_bytes, rslo_stream_t stream)
{
    if (!rslo_conv2d_tc_supported(Cin, Cout, ksize, stride)) {
        set_last_error("rslo_conv2d_tc_backward_data: unsupported shape", cudaErrorInvalidValue);
        return (int)cudaErrorInvalidValue;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int pad = ksize / 2;
    const int Ho = (H + 2 * pad - ksize) / stride + 1, Wo = (W + 2 * pad - ksize) / stride + 1;
    ConvTap taps[CV_MAX_TAPS];
    if (stride == 1) {
        int n = 0;
        for (int ky = 0; ky < ksize; ++ky)
            for (int kx = 0; kx < ksize; ++kx) taps[n++] = ConvTap{0, pad - kx, 0, pad - ky, ky * ksize + kx};
        return run_conv(g_split, B, Ho, Wo, Cout, 1, image_t, ksize * ksize, Cin, H, W, taps, n, dx, H, W, 1, 0, 1, 0, Cin,
                        nullptr, 0, workspace, workspace_bytes, st, nullptr, 1, accumulate);
    }
    if ((H & 1) || (W & 1)) {
        set_last_error("rslo_conv2d_tc_backward_data: stride 2 needs even H, W", cudaErrorInvalidValue);
        return (int)cudaErrorInvalidValue;
    }
    // 1x1 / stride 2: only the even-even class receives gradient, the rest of dx is zero
    if (ksize == 1 && !accumulate) RSLO_CHECK(cudaMemsetAsync(dx, 0, (size_t)B * H * W * Cin * 4, st));
    for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
            int n = 0;
            for (int ky = 0; ky < ksize; ++ky) {
                const int ey = py + pad - ky;
                if (ey & 1) continue;
                for (int kx = 0; kx < ksize; ++kx) {
                    const int ex = px + pad - kx;
                    if (ex & 1) continue;
                    taps[n++] = ConvTap{0, ex / 2, 0, ey / 2, ky * ksize + kx};
                }
            }
            if (n == 0) continue;
            const int rc = run_conv(g_split, B, Ho, Wo, Cout, 1, image_t, ksize * ksize, Cin, H / 2, W / 2, taps, n, dx, H, W, 2,
                                    py, 2, px, Cin, nullptr, 0, workspace, workspace_bytes, st, nullptr, 1, accumulate);
            if (rc) return rc;
        }
    return 0;
}

extern "C" size_t rslo_conv2d_tc_wgrad_scratch_bytes(int Cin, int Cout, int ksize)
{
    return (size_t)ksize * ksize * Cin * Cout * 4;
}

// grad_weight_oihw [Cout][Cin][k][k] (= or += with accumulate) from x_split [2][B][H][W][Cin], g_split [2][B][Ho][Wo][Cout];
// scratch: rslo_conv2d_tc_wgrad_scratch_bytes
extern "C" int rslo_conv2d_tc_backward_weight(const float* x_split, const float* g_split, int B, int H, int W, int Cin, int Cout,
                                              int ksize, int stride, int Cout_real, float* scratch, int accumulate,
                                              float* grad_weight_oihw, rslo_stream_t stream)
{
    if (!rslo_conv2d_tc_supported(Cin, Cout, ksize, stride) || (H % stride) || (W % stride)) {
        set_last_error("rslo_conv2d_tc_backward_weight: unsupported shape", cudaErrorInvalidValue);
        return (int)cudaErrorInvalidValue;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int pad = ksize / 2;
    const int Ho = (H + 2 * pad - ksize) / stride + 1, Wo = (W + 2 * pad - ksize) / stride + 1;
    WgradParams P;
    memset(&P, 0, sizeof P);
    P.ntaps = forward_taps(ksize, stride, Cin, P.taps);
    P.cpt = Cin / 32;
    P.nchunks = P.ntaps * P.cpt;
    P.nB = B;
    pick_tile(Wo, Ho, WG_PIX, &P.Wt, &P.Ht);
    P.tiles_w = cdiv(Wo, P.Wt);
    P.tiles_h = cdiv(Ho, P.Ht);
    P.steps_total = B * P.tiles_w * P.tiles_h;
    P.ngroups = cdiv(P.nchunks, 4);
    P.Cin = Cin;
    P.Cout = Cout;
    const int NT = Cout % 128 == 0 ? 128 : (Cout % 64 == 0 ? 64 : 32);
    const int ntiles = Cout / NT;
    const int maxg = 512 / NT;
    // groups per CTA: as many as TMEM holds while the grid still has >= ~148 CTAs with >= 2 pixel steps each
    int gpc = maxg < P.ngroups ? maxg : P.ngroups;
    while (gpc > 1 && (long)cdiv(P.ngroups, gpc) * ntiles * (P.steps_total / 2 > 0 ? P.steps_total / 2 : 1) < 148) --gpc;
    const int gsets = cdiv(P.ngroups, gpc);
    gpc = cdiv(P.ngroups, gsets);
    int xs = 148 / (gsets * ntiles);
    if (xs < 1) xs = 1;
    if (xs > P.steps_total) xs = P.steps_total;
    P.steps_per_cta = cdiv(P.steps_total, xs);
    xs = cdiv(P.steps_total, P.steps_per_cta);
    P.groups_per_cta = gpc;

    CUtensorMap tmX, tmG;
    int rc = encode_act(&tmX, x_split, B, H, W, Cin, stride, P.Wt, P.Ht, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;
    rc = encode_act(&tmG, g_split, B, Ho, Wo, Cout, 1, P.Wt, P.Ht, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;
    if (grad_weight_oihw != nullptr) RSLO_CHECK(cudaMemsetAsync(scratch, 0, (size_t)P.ntaps * Cin * Cout * 4, st));
    const dim3 grid(xs, gsets, ntiles);
    if (NT == 128) rc = launch_wgrad<128>(tmX, tmG, P, grid, scratch, st);
    else if (NT == 64) rc = launch_wgrad<64>(tmX, tmG, P, grid, scratch, st);
    else rc = launch_wgrad<32>(tmX, tmG, P, grid, scratch, st);
    if (rc) return rc;
    if (grad_weight_oihw == nullptr) return 0;          // left in scratch for rslo_conv2d_multi_wgrad_finish
    if (Cout_real <= 0 || Cout_real > Cout) Cout_real = Cout;
    const int total = P.ntaps * Cin * Cout_real;
    RSLO_COUNT();
    launch_pdl(k_wgrad_finish, cdiv(total, 256), 256, 0, st, scratch, Cout_real, Cout, Cin, P.ntaps, accumulate, grad_weight_oihw);
    RSLO_CHECK_LAUNCH("rslo_conv2d_tc_backward_weight(finish)");
    return 0;
}
