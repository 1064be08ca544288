// ROI threshold of the consistency loss: the k-th smallest nearest-neighbour distance, floored
// (rslo/core/losses.py:326-334: torch.kthvalue(dist, k) then max(., 1.0)).  torch's kthvalue costs
// ~110 us and a host synchronisation per call; this is a single-CTA 4-pass radix select (8 bits per
// pass over an order-preserving integer key) that leaves the threshold on the device.
#include "common.cuh"

namespace rslo {
namespace {

__device__ __forceinline__ unsigned f2key(float f)
{
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);      // total order of floats as unsigned ints
}
__device__ __forceinline__ float key2f(unsigned k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void __launch_bounds__(1024)
k_kth_threshold(const float* __restrict__ d, int n, int k, float floor_value, float* __restrict__ out)
{
    __shared__ unsigned hist[256];
    __shared__ unsigned s_prefix, s_mask;
    __shared__ int s_k;
    if (threadIdx.x == 0) { s_prefix = 0; s_mask = 0; s_k = k; }
    for (int pass = 3; pass >= 0; --pass) {
        const int shift = pass * 8;
        if (threadIdx.x < 256) hist[threadIdx.x] = 0;
        __syncthreads();
        const unsigned prefix = s_prefix, mask = s_mask;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const unsigned u = f2key(d[i]);
            if ((u & mask) == prefix) atomicAdd(&hist[(u >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int kk = s_k;
            unsigned b = 0;
            for (; b < 256; ++b) {
                const int c = (int)hist[b];
                if (kk <= c) break;
                kk -= c;
            }
            if (b > 255) b = 255;
            s_k = kk;
            s_prefix = prefix | (b << shift);
            s_mask = mask | (255u << shift);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = fmaxf(key2f(s_prefix), floor_value);
}

}  // namespace
}  // namespace rslo

using namespace rslo;

extern "C" int rslo_kth_threshold(const float* values, int n, int k, float floor_value, float* out,
                                  rslo_stream_t stream)
{
    if (n <= 0 || k < 1 || k > n) {
        set_last_error("rslo_kth_threshold: need 1 <= k <= n", cudaErrorInvalidValue);
        return (int)cudaErrorInvalidValue;
    }
    RSLO_COUNT();
    k_kth_threshold<<<1, 1024, 0, (cudaStream_t)stream>>>(values, n, k, floor_value, out);
    RSLO_CHECK_LAUNCH("rslo_kth_threshold");
    return 0;
}
