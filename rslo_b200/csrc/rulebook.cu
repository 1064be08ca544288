// K3: sparse-conv index generation (sm_100a).
//
// Replaces spconv's get_indice_pairs (dense int grid of batch*prod(spatial) cells + thrust
// sort/unique per call).  A level's active sites are held as a bitmap over its cell grid plus an
// exclusive popcount prefix (uint2 per 32 cells): membership AND the site's rank in sorted cell
// order come from ONE 8-byte load, with no hashing and no collisions.  For strided convolutions the
// sorted-unique output set the reference obtains with a sort is simply the set bits of the output
// bitmap in order, so output numbering is deterministic and bit-exact by construction.
// Tables are output-stationary (nbr[o,k] = input row or -1), so the convolution needs no atomics.
#include "common.cuh"

namespace rslo {
namespace {

__global__ void k_site_mark(const int* __restrict__ coors, int stride, int n_cap, const int* n_dev,
                            int H, int W, uint2* __restrict__ cells)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dev_count(n_dev, n_cap)) return;
    const int* c = coors + (size_t)i * stride + (stride - 3);
    unsigned key = ((unsigned)c[0] * H + c[1]) * W + c[2];
    atomicOr(&cells[key >> 5].x, 1u << (key & 31));
}

__global__ void k_site_perm(const int* __restrict__ coors, int stride, int n_cap, const int* n_dev,
                            int H, int W, const uint2* __restrict__ cells, int* __restrict__ perm)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dev_count(n_dev, n_cap)) return;
    const int* c = coors + (size_t)i * stride + (stride - 3);
    unsigned key = ((unsigned)c[0] * H + c[1]) * W + c[2];
    perm[site_lookup(cells, nullptr, key)] = i;
}

__global__ void k_subm_table(const int* __restrict__ coors, int stride, int n_cap, const int* n_dev,
                             int D, int H, int W, const uint2* __restrict__ cells,
                             const int* __restrict__ perm, int kd, int kh, int kw,
                             int* __restrict__ nbr)
{
    const int K = kd * kh * kw;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int o = (int)(t / K), k = (int)(t % K);
    if (o >= dev_count(n_dev, n_cap)) return;
    const int* c = coors + (size_t)o * stride + (stride - 3);
    int kz = k / (kh * kw), ky = (k / kw) % kh, kx = k % kw;
    int z = c[0] + kz - kd / 2, y = c[1] + ky - kh / 2, x = c[2] + kx - kw / 2;
    int v = -1;
    if (z >= 0 && z < D && y >= 0 && y < H && x >= 0 && x < W)
        v = site_lookup(cells, perm, ((unsigned)z * H + y) * W + x);
    nbr[t] = v;
}

struct ConvGeom {
    int kd, kh, kw, sd, sh, sw, pd, ph, pw, oD, oH, oW;
};

// output cell reached from input (z,y,x) through offset k, or -1
__device__ __forceinline__ long long out_cell(const int* c, int k, const ConvGeom& g)
{
    int kz = k / (g.kh * g.kw), ky = (k / g.kw) % g.kh, kx = k % g.kw;
    int z = c[0] + g.pd - kz, y = c[1] + g.ph - ky, x = c[2] + g.pw - kx;
    if (z < 0 || y < 0 || x < 0) return -1;
    if (z % g.sd || y % g.sh || x % g.sw) return -1;
    z /= g.sd; y /= g.sh; x /= g.sw;
    if (z >= g.oD || y >= g.oH || x >= g.oW) return -1;
    return ((long long)z * g.oH + y) * g.oW + x;
}

__global__ void k_strided_mark(const int* __restrict__ coors, int stride, int n_cap, const int* n_dev,
                               ConvGeom g, uint2* __restrict__ out_cells)
{
    const int K = g.kd * g.kh * g.kw;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int i = (int)(t / K), k = (int)(t % K);
    if (i >= dev_count(n_dev, n_cap)) return;
    long long oc = out_cell(coors + (size_t)i * stride + (stride - 3), k, g);
    if (oc >= 0) atomicOr(&out_cells[oc >> 5].x, 1u << (oc & 31));
}

// enumerate set bits in order -> out_coors rows (b,z,y,x)
__global__ void k_strided_coords(const uint2* __restrict__ out_cells, int nwords, ConvGeom g,
                                 int out_cap, int* __restrict__ out_coors)
{
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nwords) return;
    uint2 c = out_cells[w];
    unsigned bits = c.x;
    int r = (int)c.y;
    while (bits) {
        int b = __ffs(bits) - 1;
        bits &= bits - 1;
        if (r < out_cap) {
            long long cell = (long long)w * 32 + b;
            int x = (int)(cell % g.oW), y = (int)((cell / g.oW) % g.oH), z = (int)(cell / ((long long)g.oW * g.oH));
            int4 row = make_int4(0, z, y, x);
            *reinterpret_cast<int4*>(out_coors + (size_t)r * 4) = row;
        }
        ++r;
    }
}

__global__ void k_fill_rows(int* __restrict__ p, int K, int n_cap, const int* n_dev, int value)
{
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long n = (long long)dev_count(n_dev, n_cap) * K;
    for (; t < n; t += (long long)gridDim.x * blockDim.x) p[t] = value;
}

__global__ void k_strided_table(const int* __restrict__ coors, int stride, int n_cap, const int* n_dev,
                                ConvGeom g, const uint2* __restrict__ out_cells, int out_cap,
                                int* __restrict__ nbr, int* __restrict__ nbr_inv)
{
    const int K = g.kd * g.kh * g.kw;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int i = (int)(t / K), k = (int)(t % K);
    if (i >= dev_count(n_dev, n_cap)) return;
    long long oc = out_cell(coors + (size_t)i * stride + (stride - 3), k, g);
    int o = -1;
    if (oc >= 0) {
        o = site_lookup(out_cells, nullptr, (unsigned)oc);
        if (o >= out_cap) o = -1;
        else nbr[(size_t)o * K + k] = i;
    }
    nbr_inv[t] = o;
}

// dst[i] = src[i] >= 0 ? src[i] + add : -1 : appends one frame's table to a multi-frame table whose
// rows of that frame start `add` rows further down
__global__ void k_table_concat(const int* __restrict__ src, long long count, int add, int* __restrict__ dst)
{
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; t < count; t += (long long)gridDim.x * blockDim.x) {
        const int v = src[t];
        dst[t] = v >= 0 ? v + add : -1;
    }
}

// the same for up to RSLO_CONCAT_MAX (source, destination) segments in ONE launch: the 12 tables x T frames of a step
struct ConcatBatch {
    rslo_concat_seg_t seg[RSLO_CONCAT_MAX];
};
__global__ void k_table_concat_multi(const __grid_constant__ ConcatBatch B)
{
    const rslo_concat_seg_t& s = B.seg[blockIdx.y];
    const int4* __restrict__ src4 = reinterpret_cast<const int4*>(s.src);
    int4* __restrict__ dst4 = reinterpret_cast<int4*>(s.dst);
    const bool vec = ((((uintptr_t)s.src) | ((uintptr_t)s.dst)) & 15) == 0;
    const long long n4 = vec ? s.count >> 2 : 0;
    const int add = s.add;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += (long long)gridDim.x * blockDim.x) {
        int4 v = __ldg(src4 + t);
        v.x = v.x >= 0 ? v.x + add : -1; v.y = v.y >= 0 ? v.y + add : -1;
        v.z = v.z >= 0 ? v.z + add : -1; v.w = v.w >= 0 ? v.w + add : -1;
        dst4[t] = v;
    }
    for (long long t = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; t < s.count;
         t += (long long)gridDim.x * blockDim.x) {
        const int v = s.src[t];
        s.dst[t] = v >= 0 ? v + add : -1;
    }
}

// n[1] = raw number of output sites, n[0] = min(raw, cap): callers detect overflow from n[1] > cap
__global__ void k_clamp_count(int* n, int cap)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) n[0] = n[1] > cap ? cap : n[1];
}

}  // namespace
}  // namespace rslo

using namespace rslo;

extern "C" size_t rslo_site_table_workspace_bytes(int D, int H, int W)
{
    size_t nwords = ((size_t)D * H * W + 31) / 32;
    return ws_round(scan_ws_ints(nwords) * sizeof(int)) + 512;
}

extern "C" int rslo_site_table_build(const int32_t* coors, int coor_stride, int n_cap,
                                     const int32_t* n_dev, int D, int H, int W, uint32_t* cells_,
                                     int32_t* perm, void* workspace, size_t workspace_bytes,
                                     rslo_stream_t stream_)
{
    cudaStream_t st = (cudaStream_t)stream_;
    uint2* cells = (uint2*)cells_;
    const size_t nwords = ((size_t)D * H * W + 31) / 32;
    Workspace ws(workspace, workspace_bytes);
    int* block_sums = ws.take<int>(scan_ws_ints(nwords));
    if (!block_sums) {
        set_last_error("rslo_site_table_build: workspace too small", cudaErrorMemoryAllocation);
        return (int)cudaErrorMemoryAllocation;
    }
    RSLO_CHECK(cudaMemsetAsync(cells, 0, nwords * sizeof(uint2), st));
    if (n_cap > 0) {
        RSLO_COUNT();
        k_site_mark<<<cdiv(n_cap, 256), 256, 0, st>>>(coors, coor_stride, n_cap, n_dev, H, W, cells);
    }
    int rc = scan_cells(cells, (int)nwords, block_sums, nullptr, st);
    if (rc) return rc;
    if (perm && n_cap > 0) {
        RSLO_COUNT();
        k_site_perm<<<cdiv(n_cap, 256), 256, 0, st>>>(coors, coor_stride, n_cap, n_dev, H, W, cells, perm);
    }
    RSLO_CHECK_LAUNCH("rslo_site_table_build");
    return 0;
}

extern "C" int rslo_subm_table(const int32_t* coors, int coor_stride, int n_cap, const int32_t* n_dev,
                               int D, int H, int W, const uint32_t* cells, const int32_t* perm, int kd,
                               int kh, int kw, int32_t* nbr, rslo_stream_t stream_)
{
    if (n_cap <= 0) return 0;
    const int K = kd * kh * kw;
    RSLO_COUNT();
    k_subm_table<<<cdiv((long long)n_cap * K, 256), 256, 0, (cudaStream_t)stream_>>>(
        coors, coor_stride, n_cap, n_dev, D, H, W, (const uint2*)cells, perm, kd, kh, kw, nbr);
    RSLO_CHECK_LAUNCH("rslo_subm_table");
    return 0;
}

extern "C" size_t rslo_strided_workspace_bytes(int oD, int oH, int oW)
{
    return rslo_site_table_workspace_bytes(oD, oH, oW);
}

extern "C" int rslo_strided_table(const int32_t* coors, int coor_stride, int n_cap, const int32_t* n_dev,
                                  int D, int H, int W, int kd, int kh, int kw, int sd, int sh, int sw,
                                  int pd, int ph, int pw, uint32_t* out_cells_, int32_t* out_coors,
                                  int out_cap, int32_t* n_out_dev, int32_t* nbr, int32_t* nbr_inv,
                                  void* workspace, size_t workspace_bytes, rslo_stream_t stream_)
{
    cudaStream_t st = (cudaStream_t)stream_;
    ConvGeom g{kd, kh, kw, sd, sh, sw, pd, ph, pw,
               (D + 2 * pd - kd) / sd + 1, (H + 2 * ph - kh) / sh + 1, (W + 2 * pw - kw) / sw + 1};
    const int K = kd * kh * kw;
    uint2* out_cells = (uint2*)out_cells_;
    const size_t nwords = ((size_t)g.oD * g.oH * g.oW + 31) / 32;
    Workspace ws(workspace, workspace_bytes);
    int* block_sums = ws.take<int>(scan_ws_ints(nwords));
    if (!block_sums) {
        set_last_error("rslo_strided_table: workspace too small", cudaErrorMemoryAllocation);
        return (int)cudaErrorMemoryAllocation;
    }
    RSLO_CHECK(cudaMemsetAsync(out_cells, 0, nwords * sizeof(uint2), st));
    if (n_cap <= 0) {
        RSLO_CHECK(cudaMemsetAsync(n_out_dev, 0, 2 * sizeof(int), st));
        return 0;
    }
    const int G = cdiv((long long)n_cap * K, 256);
    RSLO_COUNT();
    k_strided_mark<<<G, 256, 0, st>>>(coors, coor_stride, n_cap, n_dev, g, out_cells);
    int rc = scan_cells(out_cells, (int)nwords, block_sums, n_out_dev + 1, st);
    if (rc) return rc;
    RSLO_COUNT();
    k_clamp_count<<<1, 32, 0, st>>>(n_out_dev, out_cap);
    RSLO_COUNT();
    k_strided_coords<<<cdiv(nwords, 256), 256, 0, st>>>(out_cells, (int)nwords, g, out_cap, out_coors);
    RSLO_COUNT();
    k_fill_rows<<<148 * 8, 256, 0, st>>>(nbr, K, out_cap, n_out_dev, -1);
    RSLO_COUNT();
    k_strided_table<<<G, 256, 0, st>>>(coors, coor_stride, n_cap, n_dev, g, out_cells, out_cap, nbr, nbr_inv);
    RSLO_CHECK_LAUNCH("rslo_strided_table");
    return 0;
}

extern "C" int rslo_table_concat(const int32_t* src, long long count, int add, int32_t* dst, rslo_stream_t stream)
{
    if (count <= 0) return 0;
    int blocks = cdiv(count, 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    RSLO_COUNT();
    k_table_concat<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, count, add, dst);
    RSLO_CHECK_LAUNCH("rslo_table_concat");
    return 0;
}

extern "C" int rslo_table_concat_multi(const rslo_concat_seg_t* segs_host, int nseg, rslo_stream_t stream)
{
    for (int first = 0; first < nseg; first += RSLO_CONCAT_MAX) {
        const int n = nseg - first < RSLO_CONCAT_MAX ? nseg - first : RSLO_CONCAT_MAX;
        ConcatBatch B;
        long long most = 1;
        for (int i = 0; i < n; ++i) {
            B.seg[i] = segs_host[first + i];
            if (B.seg[i].count > most) most = B.seg[i].count;
        }
        int blocks = cdiv(most, 256 * 8);
        if (blocks < 1) blocks = 1;
        if (blocks * n > 148 * 16) blocks = cdiv(148 * 16, n);
        RSLO_COUNT();
        k_table_concat_multi<<<dim3(blocks, n), 256, 0, (cudaStream_t)stream>>>(B);
        RSLO_CHECK_LAUNCH("rslo_table_concat_multi");
    }
    return 0;
}
