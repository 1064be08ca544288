// Shared device utilities for the rslo_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/rslo_b200.h"

namespace rslo {

void set_last_error(const char* what, cudaError_t e);

// Number of kernels this library has launched in the calling process (bench.py's gpu_launches).
extern unsigned long long g_launch_count;
#define RSLO_COUNT() (++::rslo::g_launch_count)

#define RSLO_CHECK_LAUNCH(what)                              \
    do {                                                     \
        cudaError_t e__ = cudaGetLastError();                \
        if (e__ != cudaSuccess) {                            \
            ::rslo::set_last_error(what, e__);               \
            return (int)e__;                                 \
        }                                                    \
    } while (0)

#define RSLO_CHECK(call)                                     \
    do {                                                     \
        cudaError_t e__ = (call);                            \
        if (e__ != cudaSuccess) {                            \
            ::rslo::set_last_error(#call, e__);              \
            return (int)e__;                                 \
        }                                                    \
    } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------------------
// The head runs ~880 kernels of 3-40 us per step from two CUDA graphs; the ~2 us between dependent kernel nodes is a
// visible share of it.  Kernels launched through launch_pdl() may be made resident while their predecessor in the
// stream is still running: every such kernel calls pdl_launch_dependents() at its top (the NEXT kernel may start to
// launch once all CTAs of this one have started) and pdl_wait() before its first access to global memory (blocks until
// the PREVIOUS grid has completed and its writes are visible).  Both are no-ops in a kernel launched the ordinary way.
// The attribute is set only while the caller says it is capturing a CUDA graph (rslo_set_graph_capture_hint);
// RSLO_PDL=0 turns it off altogether.
bool pdl_enabled();            // RSLO_PDL != 0 and the capture hint is set
void set_capture_hint(bool on);
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args&&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    // only between rslo_set_graph_capture_hint(1) and (0), i.e. while the head's two CUDA graphs are captured: on an
    // eager, host-bound path (eval forward) the attribute costs ~1 us of driver time per launch (362 -> 315 pairs/s),
    // and so does asking the driver with cudaStreamIsCapturing
    const bool on = pdl_enabled();
    attr[0].val.programmaticStreamSerializationAllowed = on ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Bump allocator over a caller-provided workspace (256 B aligned slices).
struct Workspace {
    char* base;
    size_t cap, off;
    __host__ Workspace(void* p, size_t bytes) : base((char*)p), cap(bytes), off(0) {}
    template <typename T>
    __host__ T* take(size_t n) {
        size_t bytes = (n * sizeof(T) + 255) & ~(size_t)255;
        if (off + bytes > cap) return nullptr;
        T* r = (T*)(base + off);
        off += bytes;
        return r;
    }
};
static inline size_t ws_round(size_t bytes) { return (bytes + 255) & ~(size_t)255; }

// n = (n_dev ? min(*n_dev, n_cap) : n_cap): every kernel that works on a data-dependent row count
// takes both, so a whole frame can be enqueued without a host round trip.
__device__ __forceinline__ int dev_count(const int* n_dev, int n_cap)
{
    if (n_dev == nullptr) return n_cap;
    int n = *n_dev;
    return n < n_cap ? n : n_cap;
}

// ---- site table: bitmap over the dense cell grid + exclusive popcount prefix ---------------------
// cells[w] = {bits, rank of the first set bit of word w}.  lookup(key) -> row index or -1.
__device__ __forceinline__ int site_lookup(const uint2* __restrict__ cells, const int* __restrict__ perm,
                                           unsigned key)
{
    uint2 w = __ldg(cells + (key >> 5));
    unsigned bit = 1u << (key & 31);
    if (!(w.x & bit)) return -1;
    int r = (int)w.y + __popc(w.x & (bit - 1));
    return perm ? __ldg(perm + r) : r;
}

// Exclusive scan of an int array (optionally through a transform) in three launches.
// SCAN_BLOCK elements per block; block_sums needs cdiv(n, SCAN_BLOCK) + 1 ints.
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_BLOCK = SCAN_THREADS * SCAN_ITEMS;

int scan_cells(uint2* cells, int nwords, int* block_sums, int* total_dev, cudaStream_t st);
int scan_ints(const int* in, int* out, int n, int* block_sums, int* total_dev, cudaStream_t st);
static inline size_t scan_ws_ints(long long n) { return (size_t)cdiv(n, SCAN_BLOCK) + 8; }

}  // namespace rslo
