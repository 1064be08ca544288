// K4-TC: sparse 3-D convolution as an implicit GEMM on the 5th-gen tensor cores (sm_100a, tcgen05).
//
//   out[o,:] = act(bias + sum_k in[nbr[o,k],:] @ W[k])        Cin, Cout in {16, 32, 64, 128, 192, 256, 512}
//
// One CTA owns 128 output rows (one TMEM lane per row) and a 64- (or 32-) wide slice of Cout (blockIdx.z).  For every kernel offset
// k that at least one of its rows uses, the producer warps gather the 128 neighbour rows (zeros
// where a row has no neighbour) into shared memory in the UMMA canonical K-major SWIZZLE_128B
// layout, while the offset's weight tile arrives by one bulk async copy (TMA unit, cp.async.bulk)
// from a pre-swizzled image.  A single elected thread issues tcgen05.mma (kind::tf32, M=128, K=8) into a TMEM
// accumulator; the drain warps read each offset's partial product back with tcgen05.ld and carry the sum over
// offsets in FP32 registers; bias + LeakyReLU epilogue.
//
// FP32 fidelity (the path's parity bound is 1e-4 relative, which plain TF32 misses): split-TF32.
// Each operand is split as x ~ hi + lo with hi = RN_tf32(x) and lo = RN_tf32(x - hi) (|lo| <= 2^-12 |x|; rounding
// lo to TF32 here, to nearest, instead of letting the tensor core truncate its low bits halves the representation
// error and removes its bias: the pair represents x to ~2^-23 relative), and
// the three products lo*hi + hi*lo + hi*hi are accumulated (lo*lo is below 2^-24 relative).  The splits
// are made while staging (A) / in the weight prep kernel (B), so the tensor core only ever sees
// operands that are already TF32-exact (no dependence on how the hardware would round).
//
// Pipeline (mbarriers): a step stages 32 input channels of one offset; full[2] (producer group + bulk-copy bytes
// -> MMA), empty[2] (tcgen05.commit -> producers), tfull[2] / tempty[2] (double-buffered TMEM accumulator: MMA <->
// drain warps).  Two 48 KB stages per CTA, two CTAs per SM.  Warp roles (544 threads, 56 registers):
//   warps 0-3 / 4-7  two producer groups, one per stage, working on alternating steps (a step is a serial chain of
//                    index lookup, gather, split, store, proxy fence, arrive: two chains in flight per CTA);
//   warp 8           TMEM owner; one elected lane (elect.sync: operands in uniform registers) issues per K block
//                    A_hi x [B_hi | B_lo] as ONE MMA of N = 2 Cout and A_lo x B_hi into the first half;
//   warps 9-16       drain + epilogue, two warps per TMEM lane quarter with half of the columns each.
// Levels whose tile count does not fill whole rounds of 296 resident CTAs are split over gridDim.y CTAs per tile
// (disjoint offsets), see tc_split_for.
#include "tc_common.cuh"

#ifndef TC_DIAG
#define TC_DIAG 0      // 1..5: timing-only diagnostics (scripts/diag_tc.sh); results are wrong when non-zero
#endif

#ifdef TC_TRACE          // per-CTA phase time stamps (scripts/tc_trace.py); never defined in the shipped build
__device__ unsigned long long g_tc_trace[8 * 4096];
#define TC_STAMP(slot)                                                                          \
    do {                                                                                        \
        if (blockIdx.x + gridDim.x * blockIdx.y < 4096) {                                       \
            unsigned long long t_;                                                              \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                             \
            g_tc_trace[8 * (blockIdx.x + gridDim.x * blockIdx.y) + (slot)] = t_;                \
        }                                                                                       \
    } while (0)
__device__ unsigned long long g_tc_steps[6 * 64];       // CTA 0: per-step stamps of each role
#define TC_STEP_STAMP(role, st)                                                                 \
    do {                                                                                        \
        if (blockIdx.x == 7 && blockIdx.y == 0 && (st) < 64) {                                  \
            unsigned long long t_;                                                              \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                             \
            g_tc_steps[(role) * 64 + (st)] = t_;                                                \
        }                                                                                       \
    } while (0)
#else
#define TC_STAMP(slot) do { } while (0)
#define TC_STEP_STAMP(role, st) do { } while (0)
#endif

namespace rslo {
namespace {
using namespace tc;

constexpr int TC_ROWS = 128;
// The gather producers pace this kernel (profiles/r02_tc_kernel_diagnostics.md: never parked, 17 % issuing, the rest
// dependency latency of one warp per scheduler), so a CTA carries 8 producer warps (16 rows each) and, to stay at two
// CTAs per SM within the register file, 8 drain warps that own half of the accumulator columns each.
constexpr int TC_PRODUCER_WARPS = 8;       // warps 0..7: gather
constexpr int TC_PRODUCERS = TC_PRODUCER_WARPS * 32;
constexpr int TC_KS = 32;                  // K channels staged per pipeline step (one 128-byte swizzle row)
constexpr int TC_STAGES = 2;
constexpr int TC_GROUP_WARPS = TC_PRODUCER_WARPS / TC_STAGES;    // producer warps per stage
constexpr int TC_MMA_WARP = TC_PRODUCER_WARPS;        // warp 8: TMEM owner and MMA issuer
constexpr int TC_DRAIN_WARPS = 8;          // warps 9..16: drain + epilogue (two per TMEM lane quarter)
constexpr int TC_DRAINERS = TC_DRAIN_WARPS * 32;
constexpr int TC_THREADS = (TC_PRODUCER_WARPS + 1 + TC_DRAIN_WARPS) * 32;     // 544

// ---- weight prep: W [K,Cin,Cout] -> per-offset images {B_hi, B_lo}, B is [NDIM x KDIM] K-major ----
//   forward      (transpose = 0): NDIM = Cout, KDIM = Cin,  B(n, kk) = W[k][kk][n]
//   data-grad    (transpose = 1): NDIM = Cin,  KDIM = Cout, B(n, kk) = W[ks][n][kk], ks = mirror ? K-1-k : k
__global__ void k_tc_prep(const float* __restrict__ W, int K, int Cin, int Cout, int transpose, int mirror,
                          int ntile, float* __restrict__ img)
{
    const int NDIM = transpose ? Cin : Cout, KDIM = transpose ? Cout : Cin;
    const int KPAD = (KDIM + 31) & ~31;                  // 16-channel inputs are zero-padded to one 32-wide block
    const int per = NDIM * KPAD;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= K * per) return;
    const int k = t / per, e = t % per, n = e / KPAD, kk = e % KPAD;
    float v = 0.f;
    if (kk < KDIM) {
        if (!transpose) v = W[((size_t)k * Cin + kk) * Cout + n];
        else v = W[((size_t)(mirror ? K - 1 - k : k) * Cin + n) * Cout + kk];
    }
    const float hi = tf32_rn(v), lo = tf32_rn(v - hi);
    // image: per N tile nt (ntile output channels), per offset k, per 32-wide K block kb:
    //        {B_hi [ntile x 32], B_lo [ntile x 32]}, each SWIZZLE_128B K-major
    const int kb = kk >> 5, nt = n / ntile, nn = n % ntile;
    char* base = (char*)img + (((size_t)nt * K + k) * (KPAD / 32) + kb) * (size_t)(2 * ntile * 128);
    const uint32_t off = sw128_offset(nn, kk & 31, ntile);
    *(float*)(base + off) = hi;
    *(float*)(base + (size_t)ntile * 128 + off) = lo;
}

template <int KDIM, int NDIM>
struct TcSmem {
    static constexpr int A_BYTES = TC_ROWS * TC_KS * 4;            // one of {hi, lo}: 16 KB
    static constexpr int B_BYTES = NDIM * TC_KS * 4;               // one of {hi, lo}
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int NBR_BYTES = TC_ROWS * 27 * 4;             // K <= 27 (3x3x3)
    static constexpr int TOTAL = TC_STAGES * STAGE_BYTES + NBR_BYTES + 256 + 1024;   // + barriers + alignment slack
};

// Named barrier among the drain threads only (barrier 0 is __syncthreads).
__device__ __forceinline__ void drain_sync() { asm volatile("bar.sync 1, %0;" ::"n"(TC_DRAINERS) : "memory"); }

// grid = (row tiles, split).  CTA (tile, sidx) handles every split-th active kernel offset of its tile;
// with split > 1 each CTA parks its partial sums in `scratch` and the last one to finish adds them in
// fixed order (deterministic), applies bias + activation and writes the rows.
template <int KDIM, int NDIM>
__global__ void __launch_bounds__(TC_THREADS, 2)
k_spconv_tc(const float* __restrict__ in, const int* __restrict__ nbr, int n_cap, const int* n_dev, int K,
            const float* __restrict__ bimg, const float* __restrict__ bias, int act, float slope,
            float* __restrict__ out, int n_total, float* __restrict__ scratch, int* __restrict__ tile_counter)
{
    using S = TcSmem<KDIM, NDIM>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    int* s_nbr = (int*)(smem + TC_STAGES * S::STAGE_BYTES);
    uint64_t* bars = (uint64_t*)(smem + TC_STAGES * S::STAGE_BYTES + S::NBR_BYTES);
    uint64_t* full_bar = bars;                          // [TC_STAGES] producers (+ weight-copy bytes) -> MMA
    uint64_t* empty_bar = bars + TC_STAGES;             // [TC_STAGES] MMA retired -> producers
    uint64_t* tfull_bar = bars + 2 * TC_STAGES;         // [2] accumulator buffer complete -> drain warps
    uint64_t* tempty_bar = bars + 2 * TC_STAGES + 2;    // [2] drained -> MMA
    uint32_t* s_tmem = (uint32_t*)(bars + 2 * TC_STAGES + 4);
    uint32_t* s_mask = s_tmem + 1;
    int* s_last = (int*)(s_tmem + 2);
    int* s_klist = (int*)(s_tmem + 3);        // [27]
    constexpr int KPAD = (KDIM + TC_KS - 1) / TC_KS * TC_KS;   // 16 real channels occupy one zero-padded 32-wide step
    constexpr int NSUB = KPAD / TC_KS;        // pipeline steps per kernel offset

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n = dev_count(n_dev, n_cap);
    const int row0 = blockIdx.x * TC_ROWS;
    const int split = gridDim.y, sidx = blockIdx.y;
    const int ntile = blockIdx.z;             // this CTA's slice of NDIM output channels (n_total = gridDim.z * NDIM)
    const int wtile = blockIdx.x * gridDim.z + ntile;          // work tile id (rows x channel slice)
    bimg += (size_t)ntile * K * NSUB * (2 * S::B_BYTES / 4);
    if (row0 >= n) return;                    // uniform per CTA (all splits of the tile agree)
    if (tid == 0) TC_STAMP(0);
    // two accumulator buffers of 2 * NDIM columns each: [A_hi B_hi + A_lo B_hi | A_hi B_lo] (64, 128 or 256 columns)
    constexpr int TCOLS = 4 * NDIM;

    if (tid == 0) {
        for (int i = 0; i < TC_STAGES; ++i) {
            mbar_init(full_bar + i, TC_GROUP_WARPS * 32);
            mbar_init(empty_bar + i, 1);
        }
        mbar_init(tfull_bar + 0, 1);
        mbar_init(tfull_bar + 1, 1);
        mbar_init(tempty_bar + 0, TC_DRAINERS);
        mbar_init(tempty_bar + 1, TC_DRAINERS);
        *s_mask = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TC_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(TCOLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    // neighbour tile -> smem, and the set of offsets this tile uses
    unsigned my_mask = 0;
#pragma unroll 4
    for (int i = tid; i < TC_ROWS * K; i += TC_THREADS) {
        const int r = i / K, k = i - r * K;
        const int o = row0 + r;
        const int v = o < n ? __ldg(nbr + (size_t)o * K + k) : -1;
        s_nbr[i] = v;
        if (v >= 0) my_mask |= 1u << k;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) my_mask |= __shfl_xor_sync(0xffffffffu, my_mask, d);
    if (lane == 0 && my_mask) atomicOr(s_mask, my_mask);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // this CTA's offsets: every split-th set bit of the tile's mask
    unsigned mask = 0;
    {
        unsigned m = *s_mask;
        for (int j = 0; m; ++j) {
            const unsigned bit = m & (0u - m);
            m ^= bit;
            if (j % split == sidx) mask |= bit;
        }
    }
    const uint32_t tmem_base = *s_tmem;
    if (tid == 0) {                           // this CTA's offsets in order, for indexed access by step
        unsigned m = mask;
        for (int j = 0; m; ++j) {
            s_klist[j] = __ffs(m) - 1;
            m &= m - 1;
        }
        TC_STAMP(1);
    }
    __syncthreads();

    if (warp < TC_PRODUCER_WARPS) {
        // ================= producers: gather neighbour rows, split to TF32 hi/lo, swizzled store =================
        constexpr int CHUNKS = TC_KS / 4;                // 16-byte chunks per staged row segment (128 B)
        constexpr int ROWS_PER_LD = 32 / CHUNKS;         // 4 rows per warp-wide load
        // One producer group per pipeline stage; the groups work on alternating steps.  A step is a serial chain per
        // warp (index lookup -> gather -> split -> store -> proxy fence -> arrive: ~1.1 us measured with every warp on
        // every step, profiles/r02_tc_kernel_diagnostics.md), so two independent chains in flight are worth more than
        // twice the threads on one.
        constexpr int ROWS_PER_WARP = TC_ROWS / TC_GROUP_WARPS;          // 32
        constexpr int NLD = ROWS_PER_WARP / ROWS_PER_LD; // 8 loads per thread per step, all in flight together
        const int grp = warp / TC_GROUP_WARPS, gw = warp % TC_GROUP_WARPS;
        const int sub = lane / CHUNKS, c = lane % CHUNKS;
        const uint32_t smem_base = smem_u32(smem);
        // per-thread swizzled store offsets of its NLD row segments (same for every step)
        uint32_t soff[NLD];
#pragma unroll
        for (int j = 0; j < NLD; ++j) soff[j] = sw128_offset(gw * ROWS_PER_WARP + j * ROWS_PER_LD + sub, c * 4, TC_ROWS);
        int nsteps = __popc(mask) * NSUB;

        // gather of step st: NLD independent 16-byte loads (zeros for rows without this neighbour)
        auto gather = [&](int k, int h, float4(&v)[NLD]) {
#pragma unroll
            for (int j = 0; j < NLD; ++j) {
                const int r = gw * ROWS_PER_WARP + j * ROWS_PER_LD + sub;
                const int src = s_nbr[r * K + k];
                v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#if TC_DIAG != 3
                if (src >= 0 && h * TC_KS + c * 4 < KDIM)
                    v[j] = __ldg(reinterpret_cast<const float4*>(in + (size_t)src * KDIM + h * TC_KS) + c);
#endif
            }
        };
        // split to TF32 {hi, lo} and store into stage s in the UMMA layout
        auto stage_store = [&](int st, int k, int h, const float4(&v)[NLD]) {
            const int s = st % TC_STAGES;
            if (gw == 0 && lane == 0) TC_STEP_STAMP(0, st);       // about to wait for the stage
            mbar_wait(empty_bar + s, ((st / TC_STAGES) & 1) ^ 1);
            if (gw == 0 && lane == 0) TC_STEP_STAMP(1, st);       // stage free
            const uint32_t stage = smem_base + s * S::STAGE_BYTES;
#if TC_DIAG == 1
            if (gw == 0 && st < TC_STAGES && elect_one()) {
#else
            if (gw == 0 && elect_one()) {        // one lane, uniform operands for the bulk copy
#endif
                mbar_expect_tx(full_bar + s, 2 * S::B_BYTES);
                bulk_copy_g2s(smem + s * S::STAGE_BYTES + 2 * S::A_BYTES,
                              (const char*)bimg + ((size_t)k * NSUB + h) * (2 * S::B_BYTES), 2 * S::B_BYTES, full_bar + s);
            }
#pragma unroll
            for (int j = 0; j < NLD; ++j) {
#if TC_DIAG == 2
                const float4 hh = v[j], ll = v[j];
#else
                const float4 hh = make_float4(tf32_rn(v[j].x), tf32_rn(v[j].y), tf32_rn(v[j].z), tf32_rn(v[j].w));
                const float4 ll = make_float4(tf32_rn(v[j].x - hh.x), tf32_rn(v[j].y - hh.y), tf32_rn(v[j].z - hh.z), tf32_rn(v[j].w - hh.w));
#endif
                sts128(stage + soff[j], hh);
                sts128(stage + S::A_BYTES + soff[j], ll);
            }
        };
        // make the stores visible to the tensor core's (async-proxy) reads and hand the stage over
        auto publish = [&](int st) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(full_bar + st % TC_STAGES);
            if (gw == 0 && lane == 0) TC_STEP_STAMP(2, st);       // published
        };
        // this group's steps: grp, grp + TC_STAGES, ... (always stage `grp`); the loads of its next step are issued
        // right after the hand-over and are in flight while the other group stores and publishes (issuing them before
        // the proxy fence measured 2-5 % slower)
        auto kof = [&](int st) { return s_klist[st / NSUB]; };
        float4 va[NLD];
        int st = grp;
        if (st < nsteps) gather(kof(st), st % NSUB, va);
        for (; st < nsteps; st += TC_STAGES) {
            stage_store(st, kof(st), st % NSUB, va);
            publish(st);
            if (st + TC_STAGES < nsteps) gather(kof(st + TC_STAGES), (st + TC_STAGES) % NSUB, va);
        }
        if (tid == 0) {
            TC_STAMP(2);
#ifdef TC_TRACE
            g_tc_trace[8 * (blockIdx.x + gridDim.x * blockIdx.y) + 6] = nsteps;
#endif
        }
    } else if (warp == TC_MMA_WARP) {
        // ================= MMA issuer (one elected lane) =================
        // `elect_one()` instead of `lane == 0`: the region is single-threaded by construction, so the compiler keeps the
        // tcgen05 operands in uniform registers; under `lane == 0` every tcgen05.mma sat inside a generated broadcast
        // loop (ELECT / R2UR.BROADCAST / BRA.U.ANY).
        if (elect_one()) {
            // The image holds {B_hi, B_lo} of a step back to back, which IS one [2 NDIM x 32] K-major tile: A_hi meets
            // both in ONE MMA of N = 2 NDIM (columns [0, NDIM) = hi*hi, [NDIM, 2 NDIM) = hi*lo), A_lo * B_hi adds into the
            // first half.  Two shared-memory passes over the gathered operand per K block instead of three.
            constexpr uint32_t idesc = umma_idesc_tf32(NDIM), idesc2 = umma_idesc_tf32(2 * NDIM);
            unsigned m = mask;
            for (int it = 0; m; ++it) {
                m &= m - 1;
                const int buf = it & 1;
                mbar_wait(tempty_bar + buf, ((it >> 1) & 1) ^ 1);     // accumulator buffer drained
                const uint32_t d = tmem_base + buf * 2 * NDIM;
                uint32_t acc = 0;                                     // each offset starts a fresh accumulation
#pragma unroll
                for (int h = 0; h < NSUB; ++h) {
                    const int st = it * NSUB + h, s = st % TC_STAGES;
                    mbar_wait(full_bar + s, (st / TC_STAGES) & 1);    // operands staged
                    TC_STEP_STAMP(3, st);                             // MMA warp saw the stage
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_hi = smem_u32(smem + s * S::STAGE_BYTES);
                    const uint32_t a_lo = a_hi + S::A_BYTES;
                    const uint32_t b_hi = a_hi + 2 * S::A_BYTES;      // B_lo follows at + S::B_BYTES
#if TC_DIAG != 4
#pragma unroll
                    for (int kk = 0; kk < TC_KS / 8; ++kk) {          // hi*hi | hi*lo   (lo*lo < 2^-24 relative)
                        umma_tf32(d, umma_desc_k_sw128(a_hi + kk * 32), umma_desc_k_sw128(b_hi + kk * 32), idesc2, acc);
                        acc = 1;
                    }
#pragma unroll
                    for (int kk = 0; kk < TC_KS / 8; ++kk)            // lo*hi
                        umma_tf32(d, umma_desc_k_sw128(a_lo + kk * 32), umma_desc_k_sw128(b_hi + kk * 32), idesc, 1u);
#endif
                    umma_commit(empty_bar + s);        // stage reusable once these MMAs retire
                    TC_STEP_STAMP(4, st);                             // MMAs issued
                }
                umma_commit(tfull_bar + buf);          // this offset's partial product is complete
            }
            TC_STAMP(3);
        }
        __syncwarp();
    } else {
        // ================= drain + epilogue warps (one output row x half of the channels per thread) ==============
        // The tensor core's accumulator adds are not round-to-nearest; only the 4*KDIM/8 MMAs of ONE
        // offset accumulate in TMEM, the sum over offsets is carried here in FP32 registers (RN adds).
        const int q = warp & 3;                           // TMEM lane quarter this warp may access (hardware: warp % 4)
        const int half = (warp - TC_MMA_WARP - 1) >> 2;   // warps 9..12 take the low columns, 13..16 the high ones
        constexpr int HC = NDIM / 2;                      // accumulator columns per thread
        const int r = q * 32 + lane;
        const int o = row0 + r;
        float acc[HC];
#pragma unroll
        for (int i = 0; i < HC; ++i) acc[i] = 0.f;
        unsigned m = mask;
        for (int it = 0; m; ++it) {
            m &= m - 1;
            const int buf = it & 1;
            mbar_wait(tfull_bar + buf, (it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 2 * NDIM + half * HC;
#if TC_DIAG != 5
#pragma unroll
            for (int cb = 0; cb < HC; cb += 8) {
                float v[8], w[8];
                tmem_ld8x2(taddr + cb, taddr + NDIM + cb, v, w);      // (hi*hi + lo*hi) and hi*lo of the same outputs
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[cb + i] += v[i] + w[i];
            }
#endif
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(tempty_bar + buf);
            if (warp == TC_MMA_WARP + 1 && lane == 0) TC_STEP_STAMP(5, it);   // offset drained
        }
        if (warp == TC_MMA_WARP + 1 && lane == 0) TC_STAMP(4);
        bool finish = true;
        if (split > 1) {
            float4* mine = reinterpret_cast<float4*>(scratch + ((size_t)(wtile * split + sidx) * TC_ROWS + r) * NDIM + half * HC);
#pragma unroll
            for (int i = 0; i < HC; i += 4) mine[i / 4] = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
            __threadfence();
            drain_sync();
            if (warp == TC_MMA_WARP + 1 && lane == 0) *s_last = atomicAdd(tile_counter + wtile, 1) == split - 1;
            drain_sync();
            finish = *s_last != 0;
            if (finish) {
                __threadfence();
#pragma unroll
                for (int i = 0; i < HC; ++i) acc[i] = 0.f;
                for (int sp = 0; sp < split; ++sp) {           // fixed order: deterministic sum
                    const float4* p = reinterpret_cast<const float4*>(
                        scratch + ((size_t)(wtile * split + sp) * TC_ROWS + r) * NDIM + half * HC);
#pragma unroll
                    for (int i = 0; i < HC; i += 4) {
                        const float4 t = __ldcg(p + i / 4);
                        acc[i] += t.x; acc[i + 1] += t.y; acc[i + 2] += t.z; acc[i + 3] += t.w;
                    }
                }
            }
        }
        if (finish && o < n) {
            float4* dst = reinterpret_cast<float4*>(out + (size_t)o * n_total + ntile * NDIM + half * HC);
#pragma unroll
            for (int i = 0; i < HC; i += 4) {
                float4 v = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
                if (bias) {
                    const float4 b = __ldg(reinterpret_cast<const float4*>(bias + ntile * NDIM + half * HC + i));
                    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
                }
                if (act == 1) {
                    v.x = v.x > 0.f ? v.x : v.x * slope;
                    v.y = v.y > 0.f ? v.y : v.y * slope;
                    v.z = v.z > 0.f ? v.z : v.z * slope;
                    v.w = v.w > 0.f ? v.w : v.w * slope;
                }
                dst[i / 4] = v;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) TC_STAMP(5);
    if (warp == TC_MMA_WARP) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TCOLS) : "memory");
    }
}

// Offsets of a tile are split over `split` CTAs when the level is too small to fill the GPU (two CTAs
// are resident per SM): the largest split in 1..4 that keeps the whole grid in one round of 296 CTAs.  (A cost model
// that also splits levels of slightly more than 296 tiles into three rounds of half tiles was tried and measured
// slower, 107 vs 97 us at 318 tiles: the few CTAs of a last partial round run alone on their SMs and finish early.)
static inline int tc_split_for(int n_cap, int ntiles)
{
    const int tiles = cdiv(n_cap, TC_ROWS) * ntiles;
    const int split = (2 * 148) / tiles;
    return split < 1 ? 1 : (split > 4 ? 4 : split);
}

template <int KDIM, int NDIM>
int launch_tc(const float* in, const int* nbr, int n_cap, const int* n_dev, int K, const float* bimg,
              const float* bias, int act, float slope, float* out, int n_total, void* workspace,
              size_t workspace_bytes, cudaStream_t st)
{
    using S = TcSmem<KDIM, NDIM>;
    static bool configured = false;
    if (!configured) {
        RSLO_CHECK(cudaFuncSetAttribute(k_spconv_tc<KDIM, NDIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
        configured = true;
    }
    const int tiles = cdiv(n_cap, TC_ROWS), ntiles = n_total / NDIM;
    const int split = tc_split_for(n_cap, ntiles);
    float* scratch = nullptr;
    int* counter = nullptr;
    if (split > 1) {
        Workspace ws(workspace, workspace_bytes);
        counter = ws.take<int>((size_t)tiles * ntiles);
        scratch = ws.take<float>((size_t)tiles * ntiles * split * TC_ROWS * NDIM);
        if (!scratch) {
            set_last_error("rslo_spconv_tc_forward: workspace too small", cudaErrorMemoryAllocation);
            return (int)cudaErrorMemoryAllocation;
        }
        RSLO_CHECK(cudaMemsetAsync(counter, 0, (size_t)tiles * ntiles * sizeof(int), st));
    }
    RSLO_COUNT();
    k_spconv_tc<KDIM, NDIM><<<dim3(tiles, split, ntiles), TC_THREADS, S::TOTAL, st>>>(
        in, nbr, n_cap, n_dev, K, bimg, bias, act, slope, out, n_total, scratch, counter);
    RSLO_CHECK_LAUNCH("rslo_spconv_tc");
    return 0;
}

// output-channel tile of a layer: 64 where it divides, else 32
static inline int tc_ntile(int ndim) { return ndim % 64 == 0 ? 64 : (ndim % 32 == 0 ? 32 : 16); }
static inline bool tc_kdim_ok(int kdim)
{
    return kdim == 16 || kdim == 32 || kdim == 64 || kdim == 128 || kdim == 192 || kdim == 256 || kdim == 512;
}

}  // namespace
}  // namespace rslo

using namespace rslo;

#ifdef TC_TRACE
extern "C" int rslo_debug_tc_trace(unsigned long long* host, int n_words)
{
    return (int)cudaMemcpyFromSymbol(host, g_tc_trace, sizeof(unsigned long long) * n_words);
}
extern "C" int rslo_debug_tc_steps(unsigned long long* host)
{
    return (int)cudaMemcpyFromSymbol(host, g_tc_steps, sizeof(unsigned long long) * 6 * 64);
}
#endif

extern "C" int rslo_spconv_tc_supported(int Cin, int Cout, int K)
{
    // both directions must be expressible: forward (kdim = Cin, ndim = Cout) and data gradient (swapped)
    return tc_kdim_ok(Cin) && tc_kdim_ok(Cout) && K >= 1 && K <= 27;
}

extern "C" size_t rslo_spconv_tc_image_bytes(int K, int Cin, int Cout)
{
    const int a = (Cin + 31) & ~31, b = (Cout + 31) & ~31;       // either dimension may be the (padded) K side
    return (size_t)K * a * b * 8;
}

extern "C" int rslo_spconv_tc_prepare(const float* weight, int K, int Cin, int Cout, int transpose, int mirror,
                                      float* image, rslo_stream_t stream)
{
    if (!rslo_spconv_tc_supported(Cin, Cout, K)) {
        set_last_error("rslo_spconv_tc_prepare: unsupported shape", cudaErrorInvalidValue);
        return (int)cudaErrorInvalidValue;
    }
    const int ndim = transpose ? Cin : Cout, kdim = transpose ? Cout : Cin;
    const int tot = K * ndim * ((kdim + 31) & ~31);
    RSLO_COUNT();
    k_tc_prep<<<cdiv(tot, 256), 256, 0, (cudaStream_t)stream>>>(weight, K, Cin, Cout, transpose, mirror, tc_ntile(ndim),
                                                               image);
    RSLO_CHECK_LAUNCH("rslo_spconv_tc_prepare");
    return 0;
}

extern "C" size_t rslo_spconv_tc_workspace_bytes(int n_out_cap, int ndim)
{
    const int n = n_out_cap > 0 ? n_out_cap : 1;
    const int nt = tc_ntile(ndim), ntiles = ndim / nt;
    const int tiles = cdiv(n, TC_ROWS) * ntiles;
    const int split = tc_split_for(n, ntiles);
    if (split <= 1) return 256;
    return ws_round((size_t)tiles * sizeof(int)) + ws_round((size_t)tiles * split * TC_ROWS * nt * sizeof(float)) + 256;
}

extern "C" int rslo_spconv_tc_forward(const float* in, const int32_t* nbr, int n_out_cap, const int32_t* n_out_dev,
                                      int K, int kdim, int ndim, const float* image, const float* bias, int act,
                                      float slope, float* out, void* workspace, size_t workspace_bytes,
                                      rslo_stream_t stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (n_out_cap <= 0) return 0;
    if (K < 1 || K > 27 || !tc_kdim_ok(kdim) || !tc_kdim_ok(ndim)) {
        set_last_error("rslo_spconv_tc_forward: unsupported (K, kdim, ndim)", cudaErrorInvalidValue);
        return (int)cudaErrorInvalidValue;
    }
    const int nt = tc_ntile(ndim);
#define RSLO_TC_CASE(KD)                                                                                          \
    if (kdim == KD) {                                                                                             \
        if (nt == 64)                                                                                             \
            return launch_tc<KD, 64>(in, nbr, n_out_cap, n_out_dev, K, image, bias, act, slope, out, ndim,       \
                                     workspace, workspace_bytes, st);                                             \
        if (nt == 32)                                                                                             \
            return launch_tc<KD, 32>(in, nbr, n_out_cap, n_out_dev, K, image, bias, act, slope, out, ndim,       \
                                     workspace, workspace_bytes, st);                                             \
        return launch_tc<KD, 16>(in, nbr, n_out_cap, n_out_dev, K, image, bias, act, slope, out, ndim, workspace, \
                                 workspace_bytes, st);                                                            \
    }
    RSLO_TC_CASE(16)
    RSLO_TC_CASE(32)
    RSLO_TC_CASE(64)
    RSLO_TC_CASE(128)
    RSLO_TC_CASE(192)
    RSLO_TC_CASE(256)
    RSLO_TC_CASE(512)
#undef RSLO_TC_CASE
    set_last_error("rslo_spconv_tc_forward: unsupported (kdim, ndim)", cudaErrorInvalidValue);
    return (int)cudaErrorInvalidValue;
}
