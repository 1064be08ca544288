// TMA (cp.async.bulk.tensor) helpers: host-side tensor-map encoding through the driver entry point
// (no link-time dependency on libcuda) and the device-side tile loads (sm_100a).
#pragma once
#include <cuda.h>

#include "tc_common.cuh"

namespace rslo {
namespace tma {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// fp32 tensor of `rank` dims (dim 0 innermost, contiguous), strides in BYTES for dims 1..rank-1;
// out-of-bounds elements of a box (negative coordinates included) read as zero.
static inline int encode_f32(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                             const uint32_t* box, CUtensorMapSwizzle swizzle)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_last_error("cuTensorMapEncodeTiled entry point unavailable", cudaErrorNotSupported);
        return (int)cudaErrorNotSupported;
    }
    cuuint64_t gd[5];
    cuuint64_t gs[4];
    cuuint32_t bx[5], es[5];
    uint64_t extent = 4;
    for (int i = 0; i < rank; ++i) {
        gd[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
        if (i > 0) gs[i - 1] = strides_bytes[i - 1];
    }
    for (int i = 1; i < rank; ++i) {
        const uint64_t e = strides_bytes[i - 1] * dims[i];
        if (e > extent) extent = e;
    }
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char msg[256];
        snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled failed (CUresult %d; rank %d dims %llu %llu %llu box %u %u %u)", (int)r,
                 rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                 (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0);
        set_last_error(msg, cudaErrorInvalidValue);
        return (int)cudaErrorInvalidValue;
    }
    // descriptors of tensors smaller than 128 KiB: same driver workaround CUTLASS applies (drivers <= 13.1)
    int drv = 0;
    if (cudaDriverGetVersion(&drv) == cudaSuccess && drv <= 13010 && extent < 131072)
        reinterpret_cast<uint64_t*>(tm)[1] &= ~(1llu << 21);
    return 0;
}

__device__ __forceinline__ void load_5d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, int c4,
                                        uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(tm)), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(tm)), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void prefetch_desc(const CUtensorMap* tm)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
// arrive + set the expected transaction bytes of the phase (one producer thread)
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}

}  // namespace tma
}  // namespace rslo
