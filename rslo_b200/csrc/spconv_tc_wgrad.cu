// K4-TC weight gradient of the sparse convolution on tcgen05 tensor cores (sm_100a).
//
//   dW[k][ci][co] = sum_o in[nbr[o,k]][ci] * g[o][co]                     Cin, Cout in {16, 32, 64}
//
// Per kernel offset this is a [Cin x rows] x [rows x Cout] product whose reduction dimension is the
// row index.  The gathered input rows [row][ci] and the gradient rows [row][co] are exactly the
// "MN-major" UMMA operand layouts (M resp. N contiguous, K = row), so both tiles are staged as they
// lie in memory, in the canonical form for 32-bit MN-major operands (SWIZZLE_128B_BASE32B): 32-element
// (128 B) row segments, 4 rows per swizzle atom (SBO = 512 B), one [rows x 32] block per 32 channels (LBO).
// MEMB = 128 / Cin offsets are stacked along M so every MMA is M = 128: lane m of the TMEM accumulator
// of group p holds dW[p*MEMB + m / Cin][m % Cin][:].
//
// A CTA owns a contiguous slab of rows and up to GPC offset groups (one TMEM accumulator of Cout columns
// each, all resident for the CTA's lifetime); per 64-row step the gradient tile is staged once and one
// gathered input tile per group.  Operands are split x = hi + lo (hi = RN_tf32(x)) and lo*hi + hi*lo +
// hi*hi is accumulated: FP32-level products.  At the end the drain warps add the accumulators into dW
// with atomics (dW zeroed by the launcher).
#include "tc_common.cuh"

namespace rslo {
namespace {
using namespace tc;

constexpr int WG_STEP = 64;                // rows per pipeline step (8 MMAs of K = 8 per split term)
constexpr int WG_PRODUCER_WARPS = 8;
constexpr int WG_PRODUCERS = WG_PRODUCER_WARPS * 32;
constexpr int WG_GROUP = WG_PRODUCERS / 2;          // two producer groups, one per A stage, on alternating items
constexpr int WG_THREADS = (WG_PRODUCER_WARPS + 1 + 4) * 32;   // + MMA warp + 4 drain warps
constexpr int WG_ASTAGES = 2;
constexpr int WG_GSTAGES = 2;

// MN-major operands of 32-bit types have ONE legal shared-memory layout, SWIZZLE_128B_BASE32B (layout
// type 1): rows of 32 MN-contiguous elements (128 B), 4 K-rows per swizzle atom (512 B), and inside a row
// the four 32-byte chunks permuted by XOR with (K-row % 4)  [CuTe: Swizzle<2,5,2> over 1024-bit x 4 atoms].
// LBO = bytes between 32-element MN blocks, SBO = bytes between 4-row K groups.
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
// byte offset of the 16-byte chunk holding elements [col, col+4) of K-row `row` in a [rows x C] operand stored
// as C/32 blocks of [rows x 32]
__device__ __forceinline__ uint32_t mn32_offset(int row, int col, int rows)
{
    const int mb = col >> 5, c16 = (col & 31) >> 2, r4 = row & 3;
    return (uint32_t)mb * (uint32_t)(rows * 128) + (uint32_t)(row >> 2) * 512u + (uint32_t)r4 * 128u +
           (uint32_t)(((c16 >> 1) ^ r4) << 5) + (uint32_t)(c16 & 1) * 16u;
}
// kind::tf32, FP32 accumulate, A and B MN-major, M = 128, N = n
__host__ __device__ constexpr uint32_t umma_idesc_tf32_mn(int n)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(128 >> 4) << 24);
}

template <int CIN, int COUT>
struct WgCfg {
    static constexpr int MEMB = 128 / CIN;                          // offsets stacked along M
    static constexpr int MAXG = 512 / COUT > 7 ? 7 : 512 / COUT;    // accumulators per CTA (TMEM columns)
    static constexpr int A_BYTES = WG_STEP * 128 * 4;               // one of {hi, lo}: [64 rows x 128 m]
    static constexpr int G_BYTES = WG_STEP * (COUT < 32 ? 32 : COUT) * 4;   // one of {hi, lo}: [64 rows x Cout], 128-B rows
    static constexpr int TOTAL = WG_ASTAGES * 2 * A_BYTES + WG_GSTAGES * 2 * G_BYTES + 256 + 1024;
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(WG_THREADS, 1)
k_spconv_tc_wgrad(const float* __restrict__ in, const float* __restrict__ g, const int* __restrict__ nbr, int n_cap,
                  const int* n_dev, int K, int rows_per_cta, int groups_per_cta, float* __restrict__ dW)
{
    using C = WgCfg<CIN, COUT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                                              // [WG_ASTAGES][hi | lo]
    uint8_t* sG = smem + WG_ASTAGES * 2 * C::A_BYTES;                // [WG_GSTAGES][hi | lo]
    uint64_t* bars = (uint64_t*)(sG + WG_GSTAGES * 2 * C::G_BYTES);
    uint64_t* a_full = bars;                  // [2] producers -> MMA
    uint64_t* a_empty = bars + 2;             // [2] MMA retired -> producers
    uint64_t* g_full = bars + 4;              // [2]
    uint64_t* g_empty = bars + 6;             // [2]
    uint64_t* acc_bar = bars + 8;             // all MMAs retired -> drain
    uint32_t* s_tmem = (uint32_t*)(bars + 9);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n = dev_count(n_dev, n_cap);
    const int row_begin = blockIdx.x * rows_per_cta;
    const int row_end = min(n, row_begin + rows_per_cta);
    if (row_begin >= n) return;
    const int ngroups_total = (K + C::MEMB - 1) / C::MEMB;
    const int g0 = blockIdx.y * groups_per_cta;
    const int ng = min(groups_per_cta, ngroups_total - g0);          // groups this CTA accumulates
    if (ng <= 0) return;
    const int nsteps = (row_end - row_begin + WG_STEP - 1) / WG_STEP;
    constexpr int TCOLS = 512;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(a_full + i, WG_GROUP);
            mbar_init(a_empty + i, 1);
            mbar_init(g_full + i, WG_GROUP);
            mbar_init(g_empty + i, 1);
        }
        mbar_init(acc_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == WG_PRODUCER_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(TCOLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *s_tmem;

    if (warp < WG_PRODUCER_WARPS) {
        // ================= producers =================
        // Two groups of 4 warps, one per A stage, work on alternating (step, group) items: an item is a serial chain
        // (neighbour index -> gathered row -> stage wait -> split -> store -> proxy fence -> arrive), so two chains in
        // flight keep the tensor pipe fed (same finding as in spconv_tc.cu).  The group that owns the first item of a
        // step also stages the step's gradient tile.
        const uint32_t sA_u = smem_u32(sA), sG_u = smem_u32(sG);
        const int grp = warp / (WG_PRODUCER_WARPS / 2), gt = tid - grp * WG_GROUP;
        const int nitems = nsteps * ng;
        for (int item = grp; item < nitems; item += 2) {
            const int st = item / ng, gi = item - st * ng;
            const int r0 = row_begin + st * WG_STEP;
            if (gi == 0) {
                // gradient tile of this step: 64 rows x Cout, once per step
                const int gs = st & 1;
                mbar_wait(g_empty + gs, ((st >> 1) & 1) ^ 1);
                const uint32_t base = sG_u + gs * 2 * C::G_BYTES;
                constexpr int CH = COUT / 4;                         // 16-byte chunks per row
                for (int i = gt; i < WG_STEP * CH; i += WG_GROUP) {
                    const int r = i / CH, c = i % CH;
                    const int o = r0 + r;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (o < row_end) v = __ldg(reinterpret_cast<const float4*>(g + (size_t)o * COUT) + c);
                    const float4 hh = make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
                    const float4 ll = make_float4(tf32_rn(v.x - hh.x), tf32_rn(v.y - hh.y), tf32_rn(v.z - hh.z), tf32_rn(v.w - hh.w));
                    const uint32_t off = mn32_offset(r, c * 4, WG_STEP);
                    sts128(base + off, hh);
                    sts128(base + C::G_BYTES + off, ll);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(g_full + gs);
            }
            {
                const int as = item & 1;                             // == grp
                const int kbase = (g0 + gi) * C::MEMB;
                // gather first (loads in flight), then wait for the stage
                constexpr int CH = 128 / 4;                          // 32 chunks per stacked row
                constexpr int PER = WG_STEP * CH / WG_GROUP;         // 16 chunks per thread
                float4 v[PER];
#pragma unroll
                for (int j = 0; j < PER; ++j) {
                    const int i = gt + j * WG_GROUP;
                    const int r = i / CH, c = i % CH;
                    const int memb = (c * 4) / CIN, cc = (c * 4) % CIN;
                    const int k = kbase + memb;
                    const int o = r0 + r;
                    int src = -1;
                    if (o < row_end && k < K) src = __ldg(nbr + (size_t)o * K + k);
                    v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (src >= 0) v[j] = __ldg(reinterpret_cast<const float4*>(in + (size_t)src * CIN + cc));
                }
                mbar_wait(a_empty + as, ((item >> 1) & 1) ^ 1);
                const uint32_t base = sA_u + as * 2 * C::A_BYTES;
#pragma unroll
                for (int j = 0; j < PER; ++j) {
                    const int i = gt + j * WG_GROUP;
                    const int r = i / CH, c = i % CH;
                    const float4 hh = make_float4(tf32_rn(v[j].x), tf32_rn(v[j].y), tf32_rn(v[j].z), tf32_rn(v[j].w));
                    const float4 ll = make_float4(tf32_rn(v[j].x - hh.x), tf32_rn(v[j].y - hh.y), tf32_rn(v[j].z - hh.z), tf32_rn(v[j].w - hh.w));
                    const uint32_t off = mn32_offset(r, c * 4, WG_STEP);
                    sts128(base + off, hh);
                    sts128(base + C::A_BYTES + off, ll);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(a_full + as);
            }
        }
    } else if (warp == WG_PRODUCER_WARPS) {
        // ================= MMA issuer =================
        if (elect_one()) {        // single-thread region by construction: operands stay in uniform registers
            constexpr uint32_t idesc = umma_idesc_tf32_mn(COUT);
            constexpr uint32_t LBO = WG_STEP * 128;                  // bytes between 32-channel blocks
            int item = 0;
            for (int st = 0; st < nsteps; ++st) {
                const int gs = st & 1;
                mbar_wait(g_full + gs, (st >> 1) & 1);
                const uint32_t gb = smem_u32(sG) + gs * 2 * C::G_BYTES;
                for (int gi = 0; gi < ng; ++gi, ++item) {
                    const int as = item & 1;
                    mbar_wait(a_full + as, (item >> 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t ab = smem_u32(sA) + as * 2 * C::A_BYTES;
                    const uint32_t d = tmem_base + gi * COUT;
#pragma unroll
                    for (int part = 0; part < 3; ++part) {           // lo*hi, hi*lo, hi*hi
                        const uint32_t a = ab + (part == 0 ? C::A_BYTES : 0);
                        const uint32_t b = gb + (part == 1 ? C::G_BYTES : 0);
#pragma unroll
                        for (int kg = 0; kg < WG_STEP / 8; ++kg) {
                            umma_tf32(d, umma_desc_mn_sw128(a + kg * 1024, LBO, 512), umma_desc_mn_sw128(b + kg * 1024, LBO, 512),
                                      idesc, (st > 0 || part > 0 || kg > 0) ? 1u : 0u);
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
        // ================= drain: accumulators -> dW (atomics) =================
        const int q = warp & 3;
        const int m = q * 32 + lane;                                 // TMEM lane = stacked (member, ci)
        const int memb = m / CIN, ci = m % CIN;
        mbar_wait(acc_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int gi = 0; gi < ng; ++gi) {
            const int k = (g0 + gi) * C::MEMB + memb;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + gi * COUT;
#pragma unroll
            for (int cb = 0; cb < COUT; cb += 16) {
                float v[16];
                tmem_ld16(taddr + cb, v);
                if (k < K) {
                    float* dst = dW + ((size_t)k * CIN + ci) * COUT + cb;
#pragma unroll
                    for (int i = 0; i < 16; i += 4)                  // 16-byte vector reductions (REDG.F32x4)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "f"(v[i]), "f"(v[i + 1]),
                                     "f"(v[i + 2]), "f"(v[i + 3])
                                     : "memory");
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == WG_PRODUCER_WARPS) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TCOLS) : "memory");
    }
}

template <int CIN, int COUT>
int launch_wgrad(const float* in, const float* g, const int* nbr, int n_cap, const int* n_dev, int K, float* dW,
                 cudaStream_t st)
{
    using C = WgCfg<CIN, COUT>;
    static bool configured = false;
    if (!configured) {
        RSLO_CHECK(cudaFuncSetAttribute(k_spconv_tc_wgrad<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::TOTAL));
        configured = true;
    }
    const int ngroups = cdiv(K, C::MEMB);
    const int ysplit = cdiv(ngroups, C::MAXG);
    const int gpc = cdiv(ngroups, ysplit);
    int xs = 148 / ysplit;                                            // one CTA per SM
    int rows_per_cta = cdiv(cdiv(n_cap, xs), WG_STEP) * WG_STEP;
    if (rows_per_cta < WG_STEP) rows_per_cta = WG_STEP;
    xs = cdiv(n_cap, rows_per_cta);
    RSLO_COUNT();
    k_spconv_tc_wgrad<CIN, COUT><<<dim3(xs, ysplit), WG_THREADS, C::TOTAL, st>>>(in, g, nbr, n_cap, n_dev, K, rows_per_cta,
                                                                               gpc, dW);
    RSLO_CHECK_LAUNCH("rslo_spconv_tc_wgrad");
    return 0;
}

}  // namespace
}  // namespace rslo

using namespace rslo;

extern "C" int rslo_spconv_tc_wgrad_supported(int Cin, int Cout)
{
    const bool ci = Cin == 16 || Cin == 32 || Cin == 64, co = Cout == 16 || Cout == 32 || Cout == 64;
    return ci && co && !(Cin == 16 && Cout == 64) && !(Cin == 64 && Cout == 16);
}

extern "C" int rslo_spconv_tc_backward_weight(const float* in, const float* grad_out, const int32_t* nbr, int n_out_cap,
                                              const int32_t* n_out_dev, int K, int Cin, int Cout, float* grad_weight,
                                              rslo_stream_t stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    RSLO_CHECK(cudaMemsetAsync(grad_weight, 0, (size_t)K * Cin * Cout * sizeof(float), st));
    if (n_out_cap <= 0) return 0;
    if (Cin == 16 && Cout == 16) return launch_wgrad<16, 16>(in, grad_out, nbr, n_out_cap, n_out_dev, K, grad_weight, st);
    if (Cin == 16 && Cout == 32) return launch_wgrad<16, 32>(in, grad_out, nbr, n_out_cap, n_out_dev, K, grad_weight, st);
    if (Cin == 32 && Cout == 16) return launch_wgrad<32, 16>(in, grad_out, nbr, n_out_cap, n_out_dev, K, grad_weight, st);
    if (Cin == 64 && Cout == 64) return launch_wgrad<64, 64>(in, grad_out, nbr, n_out_cap, n_out_dev, K, grad_weight, st);
    if (Cin == 64 && Cout == 32) return launch_wgrad<64, 32>(in, grad_out, nbr, n_out_cap, n_out_dev, K, grad_weight, st);
    if (Cin == 32 && Cout == 64) return launch_wgrad<32, 64>(in, grad_out, nbr, n_out_cap, n_out_dev, K, grad_weight, st);
    if (Cin == 32 && Cout == 32) return launch_wgrad<32, 32>(in, grad_out, nbr, n_out_cap, n_out_dev, K, grad_weight, st);
    set_last_error("rslo_spconv_tc_backward_weight: unsupported (Cin, Cout)", cudaErrorInvalidValue);
    return (int)cudaErrorInvalidValue;
}
