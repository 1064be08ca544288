// Device-wide exclusive scans used by the voxeliser and the rulebook builder (sm_100a).
// Three launches (block reduce, top-level scan in one CTA, block down-sweep); every array here is
// at most a few MB and L2 resident, so the second read in the down-sweep never reaches HBM.
#include "common.cuh"

namespace rslo {

namespace {

__device__ __forceinline__ int warp_incl_scan(int v)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// Exclusive scan of one value per thread across the CTA; returns the exclusive prefix, *total = sum.
template <int THREADS>
__device__ __forceinline__ int block_excl_scan(int v, int* total)
{
    __shared__ int warp_sums[THREADS / 32];
    __shared__ int s_total;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = warp_incl_scan(v);
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int w = lane < THREADS / 32 ? warp_sums[lane] : 0;
        int wi = warp_incl_scan(w);
        if (lane < THREADS / 32) warp_sums[lane] = wi - w;
        if (lane == 31) s_total = wi;
    }
    __syncthreads();
    int r = incl - v + warp_sums[wid];
    *total = s_total;
    __syncthreads();
    return r;
}

struct LoadCells {
    const uint2* p;
    __device__ int operator()(int i) const { return __popc(p[i].x); }
};
struct StoreCells {
    uint2* p;
    __device__ void operator()(int i, int v) const { p[i].y = (unsigned)v; }
};
struct LoadInts {
    const int* p;
    __device__ int operator()(int i) const { return p[i]; }
};
struct StoreInts {
    int* p;
    __device__ void operator()(int i, int v) const { p[i] = v; }
};

template <typename Load>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(Load ld, int n, int* __restrict__ block_sums)
{
    const int base = blockIdx.x * SCAN_BLOCK;
    int s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        int g = base + j * SCAN_THREADS + threadIdx.x;
        if (g < n) s += ld(g);
    }
    int total;
    block_excl_scan<SCAN_THREADS>(s, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// One CTA: in-place exclusive scan of block_sums[0..nb), grand total -> block_sums[nb] and *total_dev.
__global__ void __launch_bounds__(1024) k_scan_top(int* __restrict__ block_sums, int nb, int* __restrict__ total_dev)
{
    int carry = 0;
    for (int base = 0; base < nb; base += 1024) {
        int i = base + threadIdx.x;
        int v = i < nb ? block_sums[i] : 0;
        int total;
        int ex = block_excl_scan<1024>(v, &total);
        if (i < nb) block_sums[i] = ex + carry;
        carry += total;
    }
    if (threadIdx.x == 0) {
        block_sums[nb] = carry;
        if (total_dev) *total_dev = carry;
    }
}

template <typename Load, typename Store>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(Load ld, Store st, int n, const int* __restrict__ block_sums)
{
    const int base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        int g = base + j;
        v[j] = g < n ? ld(g) : 0;
        s += v[j];
    }
    int total;
    int ex = block_excl_scan<SCAN_THREADS>(s, &total) + block_sums[blockIdx.x];
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        int g = base + j;
        if (g < n) st(g, ex);
        ex += v[j];
    }
}

template <typename Load, typename Store>
int scan_impl(Load ld, Store st, int n, int* block_sums, int* total_dev, cudaStream_t stream)
{
    if (n <= 0) {
        if (total_dev) RSLO_CHECK(cudaMemsetAsync(total_dev, 0, sizeof(int), stream));
        return 0;
    }
    const int nb = cdiv(n, SCAN_BLOCK);
    RSLO_COUNT();
    k_scan_reduce<<<nb, SCAN_THREADS, 0, stream>>>(ld, n, block_sums);
    RSLO_COUNT();
    k_scan_top<<<1, 1024, 0, stream>>>(block_sums, nb, total_dev);
    RSLO_COUNT();
    k_scan_apply<<<nb, SCAN_THREADS, 0, stream>>>(ld, st, n, block_sums);
    RSLO_CHECK_LAUNCH("scan");
    return 0;
}

}  // namespace

int scan_cells(uint2* cells, int nwords, int* block_sums, int* total_dev, cudaStream_t st)
{
    return scan_impl(LoadCells{cells}, StoreCells{cells}, nwords, block_sums, total_dev, st);
}

int scan_ints(const int* in, int* out, int n, int* block_sums, int* total_dev, cudaStream_t st)
{
    return scan_impl(LoadInts{in}, StoreInts{out}, n, block_sums, total_dev, st);
}

}  // namespace rslo
