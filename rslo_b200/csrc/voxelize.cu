// K2: scatter-voxeliser with fused VFE mean (sm_100a).
//
// Reproduces, in parallel, the exact output of the sequential first-come scan the reference runs on
// the CPU (spconv VoxelGenerator.generate, called at rslo/data/preprocess.py:493):
//   * voxel ids in order of each voxel's first point,
//   * the first `max_points` points of each voxel, in point order,
//   * only the first `max_voxels` voxels,
//   * optional block height filter, compaction order preserving.
// Method: a bitmap over the cell grid + popcount prefix gives every occupied cell a dense rank r
// (sorted cell order) with no hash collisions; atomicMin over point indices finds each cell's first
// point; a second bitmap over POINT indices ranks those first points, which is the first-come voxel
// id; a counting sort groups the point indices per cell and one thread per voxel selects its
// `max_points` smallest indices.  Everything is integer-exact and deterministic.
// HBM traffic: points are read once as keys (12 B/pt) and once more for the gathered rows; all
// scratch (bitmap 5.5 MB + prefix, per-point ints) lives in L2.
#include <limits.h>

#include "common.cuh"

namespace rslo {
namespace {

constexpr unsigned INVALID_KEY = 0xffffffffu;
constexpr int EMPTY_MIN = 0x7f7f7f7f;        // memset(0x7f): above every finite ordered-float key
constexpr int EMPTY_MAX = (int)0x80808080u;  // memset(0x80): below every finite ordered-float key

struct VoxParams {
    float vx, vy, vz, x0, y0, z0;
    int gx, gy, gz;
};

__global__ void k_vox_mark(const float* __restrict__ pts, int P, int F, VoxParams vp,
                           unsigned* __restrict__ keys, uint2* __restrict__ cells)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float* p = pts + (size_t)i * F;
    // float32 divide + floor, as the reference (IEEE division; no fast-math in this build)
    int cx = (int)floorf((p[0] - vp.x0) / vp.vx);
    int cy = (int)floorf((p[1] - vp.y0) / vp.vy);
    int cz = (int)floorf((p[2] - vp.z0) / vp.vz);
    unsigned key = INVALID_KEY;
    if (cx >= 0 && cx < vp.gx && cy >= 0 && cy < vp.gy && cz >= 0 && cz < vp.gz) {
        key = ((unsigned)cz * vp.gy + cy) * vp.gx + cx;
        atomicOr(&cells[key >> 5].x, 1u << (key & 31));
    }
    keys[i] = key;
}

__global__ void k_vox_first(const unsigned* __restrict__ keys, int P, const uint2* __restrict__ cells,
                            int* __restrict__ ranks, int* __restrict__ first, int* __restrict__ cnt)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    unsigned key = keys[i];
    int r = -1;
    if (key != INVALID_KEY) {
        r = site_lookup(cells, nullptr, key);
        atomicMin(first + r, i);
        atomicAdd(cnt + r, 1);
    }
    ranks[i] = r;
}

__global__ void k_vox_mark_first(const int* __restrict__ first, const int* __restrict__ n_cells_dev,
                                 uint2* __restrict__ ptbits)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= *n_cells_dev) return;
    int f = first[r];
    atomicOr(&ptbits[f >> 5].x, 1u << (f & 31));
}

// vid[r] = first-come voxel id of cell rank r; also the (optional) block min/max of z.
__global__ void k_vox_scatter(const float* __restrict__ pts, int P, int F, const int* __restrict__ ranks,
                              const int* __restrict__ first, const uint2* __restrict__ ptbits,
                              const int* __restrict__ offsets, int* __restrict__ fill,
                              int* __restrict__ seg, int max_voxels, const unsigned* __restrict__ keys,
                              int gx, int gy, int block_factor, int bw, int bh,
                              int* __restrict__ zmin, int* __restrict__ zmax)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    int r = ranks[i];
    if (r < 0) return;
    int vid = site_lookup(ptbits, nullptr, (unsigned)first[r]);
    if (vid >= max_voxels) return;
    int pos = offsets[r] + atomicAdd(fill + r, 1);
    seg[pos] = i;
    if (zmin) {
        unsigned key = keys[i];
        int x = key % gx, y = (key / gx) % gy;
        int bx = x / block_factor, by = y / block_factor;
        if (bx < bw && by < bh) {
            // order-preserving float -> int map so integer atomics give float min/max
            int zi = __float_as_int(pts[(size_t)i * F + 2]);
            zi = zi >= 0 ? zi : zi ^ 0x7fffffff;
            atomicMin(zmin + by * bw + bx, zi);
            atomicMax(zmax + by * bw + bx, zi);
        }
    }
}

__device__ __forceinline__ float ordered_int_to_float(int zi)
{
    return __int_as_float(zi >= 0 ? zi : zi ^ 0x7fffffff);
}

// keep[vid] = 1 iff the voxel passes the block height filter.
__global__ void k_vox_filter(const int* __restrict__ n_cells_dev, const int* __restrict__ first,
                             const uint2* __restrict__ ptbits, const unsigned* __restrict__ keys,
                             int max_voxels, int gx, int gy, int block_factor, int block_size, int bw,
                             int bh, const int* __restrict__ zmin, const int* __restrict__ zmax,
                             float thr, int* __restrict__ keep)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= *n_cells_dev) return;
    int f = first[r];
    int vid = site_lookup(ptbits, nullptr, (unsigned)f);
    if (vid >= max_voxels) return;
    unsigned key = keys[f];
    int x = key % gx, y = (key / gx) % gy;
    int bx = x / block_factor, by = y / block_factor;
    int y0 = max(by - block_size / 2, 0), y1 = min(by + block_size - block_size / 2, bh);
    int x0 = max(bx - block_size / 2, 0), x1 = min(bx + block_size - block_size / 2, bw);
    int mn = EMPTY_MIN, mx = EMPTY_MAX;
    for (int yy = y0; yy < y1; ++yy)
        for (int xx = x0; xx < x1; ++xx) {
            mn = min(mn, zmin[yy * bw + xx]);
            mx = max(mx, zmax[yy * bw + xx]);
        }
    float d = (mx == EMPTY_MAX || mn == EMPTY_MIN) ? -INFINITY
                                               : ordered_int_to_float(mx) - ordered_int_to_float(mn);
    // empty windows: the reference compares (-inf) - (+inf) = -inf against the threshold
    keep[vid] = d > thr ? 1 : 0;
}

constexpr int MAXP_LIMIT = 32;

// One thread per occupied cell.  The `max_points` smallest point indices of the cell are kept in a
// register-resident sorted list (static-index compare-exchange insertion: MAXP is a compile-time bound, so
// nothing spills to local memory), then the selected rows are read once for `voxels` and/or the fused mean.
template <int MAXP>
__global__ void __launch_bounds__(128)
k_vox_gather(const float* __restrict__ pts, int F, const int* __restrict__ n_cells_dev,
             const int* __restrict__ first, const uint2* __restrict__ ptbits,
             const int* __restrict__ offsets, const int* __restrict__ cnt,
             const int* __restrict__ seg, const unsigned* __restrict__ keys,
             int max_points, int max_voxels, int gx, int gy,
             const int* __restrict__ keep, const int* __restrict__ newid, int batch_idx,
             float* __restrict__ voxels, int* __restrict__ coors, int coor_stride,
             int* __restrict__ num_points, float* __restrict__ mean,
             int* __restrict__ perm)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= *n_cells_dev) return;
    int f = first[r];
    int vid = site_lookup(ptbits, nullptr, (unsigned)f);
    int row = vid;
    if (vid >= max_voxels) row = -1;
    else if (keep) row = keep[vid] ? newid[vid] : -1;
    if (perm) perm[r] = row;
    if (row < 0) return;

    int sel[MAXP];
#pragma unroll
    for (int q = 0; q < MAXP; ++q) sel[q] = INT_MAX;
    const int c = cnt[r];
    const int* s = seg + offsets[r];
    for (int j = 0; j < c; ++j) {
        int v = __ldg(s + j);
#pragma unroll
        for (int q = 0; q < MAXP; ++q) {                 // keeps sel[] ascending; v carries the displaced value
            const int lo = min(v, sel[q]);
            v = max(v, sel[q]);
            sel[q] = lo;
        }
    }
    const int m = min(c, max_points);
    unsigned key = keys[f];
    int x = key % gx, y = (key / gx) % gy, z = key / (gx * gy);
    int* co = coors + (size_t)row * coor_stride;
    if (coor_stride == 4) { co[0] = batch_idx; ++co; }
    co[0] = z; co[1] = y; co[2] = x;
    num_points[row] = m;
    float acc[7] = {0, 0, 0, 0, 0, 0, 0};
    float* vo = voxels ? voxels + (size_t)row * max_points * F : nullptr;
#pragma unroll
    for (int q = 0; q < MAXP; ++q) {
        if (q < max_points) {
            if (q < m) {
                const float* p = pts + (size_t)sel[q] * F;
                if (vo)
                    for (int a = 0; a < F; ++a) vo[q * F + a] = p[a];
                if (mean) {
#pragma unroll
                    for (int a = 0; a < 7; ++a) acc[a] += p[a];
                }
            } else if (vo) {
                for (int a = 0; a < F; ++a) vo[q * F + a] = 0.f;
            }
        }
    }
    if (mean) {
        // SimpleVoxel_XYZINormalC (voxel_encoder.py:272-280): mean of the stored points, then the
        // normal (cols 4:7) renormalised with eps 1e-12.
        float fm = (float)m;
#pragma unroll
        for (int a = 0; a < 7; ++a) acc[a] = acc[a] / fm;
        float nrm = sqrtf(acc[4] * acc[4] + acc[5] * acc[5] + acc[6] * acc[6]) + 1e-12f;
        float* mo = mean + (size_t)row * 7;
        mo[0] = acc[0]; mo[1] = acc[1]; mo[2] = acc[2]; mo[3] = acc[3];
        mo[4] = acc[4] / nrm; mo[5] = acc[5] / nrm; mo[6] = acc[6] / nrm;
    }
}

__global__ void k_vox_count(const int* __restrict__ n_cells_dev, int max_voxels,
                            const int* __restrict__ kept_total, int* __restrict__ n_out)
{
    if (threadIdx.x == 0 && blockIdx.x == 0)
        *n_out = kept_total ? *kept_total : min(*n_cells_dev, max_voxels);
}

__global__ void k_vfe_mean(const float* __restrict__ voxels, const int* __restrict__ num_points, int n,
                           int max_points, int F, float* __restrict__ mean)
{
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const float* p = voxels + (size_t)v * max_points * F;
    float acc[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int q = 0; q < max_points; ++q)
#pragma unroll
        for (int a = 0; a < 7; ++a) acc[a] += p[q * F + a];
    float fm = (float)num_points[v];
#pragma unroll
    for (int a = 0; a < 7; ++a) acc[a] = acc[a] / fm;
    float nrm = sqrtf(acc[4] * acc[4] + acc[5] * acc[5] + acc[6] * acc[6]) + 1e-12f;
    float* mo = mean + (size_t)v * 7;
    mo[0] = acc[0]; mo[1] = acc[1]; mo[2] = acc[2]; mo[3] = acc[3];
    mo[4] = acc[4] / nrm; mo[5] = acc[5] / nrm; mo[6] = acc[6] / nrm;
}

}  // namespace
}  // namespace rslo

using namespace rslo;

extern "C" size_t rslo_voxelize_workspace_bytes(int P, int gx, int gy, int table_d, int max_voxels)
{
    size_t nwords = ((size_t)table_d * gy * gx + 31) / 32;
    size_t pwords = ((size_t)P + 31) / 32;
    size_t b = 0;
    b += ws_round(nwords * sizeof(uint2));                 // cells (when the caller gives none)
    b += ws_round(pwords * sizeof(uint2));                 // ptbits
    b += ws_round(scan_ws_ints(nwords > (size_t)P ? nwords : P) * sizeof(int));
    b += 7 * ws_round((size_t)P * sizeof(int));            // keys ranks first cnt offsets fill seg
    b += 2 * ws_round((size_t)max_voxels * sizeof(int));   // keep newid
    b += 2 * ws_round((size_t)gx * gy * sizeof(int));      // zmin zmax
    b += 4 * 256;
    return b;
}

extern "C" int rslo_voxelize(const float* points, int P, int F, const float* vs, const float* rg,
                             int gx, int gy, int gz, int max_points, int max_voxels, int block_factor,
                             int block_size, float height_threshold, int batch_idx, float* voxels,
                             int32_t* coors, int coor_stride, int32_t* num_points, float* mean,
                             int32_t* n_voxels_dev, uint32_t* cells_out, int32_t* perm, int table_d,
                             void* workspace, size_t workspace_bytes, rslo_stream_t stream_)
{
    cudaStream_t st = (cudaStream_t)stream_;
    if (max_points > MAXP_LIMIT || max_points < 1 || (mean && F < 7) || F < 3 ||
        (coor_stride != 3 && coor_stride != 4) || table_d < gz) {
        set_last_error("rslo_voxelize: bad argument", cudaErrorInvalidValue);
        return (int)cudaErrorInvalidValue;
    }
    if (P <= 0) {
        RSLO_CHECK(cudaMemsetAsync(n_voxels_dev, 0, sizeof(int), st));
        return 0;
    }
    Workspace ws(workspace, workspace_bytes);
    const size_t nwords = ((size_t)table_d * gy * gx + 31) / 32;
    const size_t pwords = ((size_t)P + 31) / 32;
    uint2* cells = cells_out ? (uint2*)cells_out : ws.take<uint2>(nwords);
    uint2* ptbits = ws.take<uint2>(pwords);
    int* block_sums = ws.take<int>(scan_ws_ints(nwords > (size_t)P ? nwords : P));
    unsigned* keys = ws.take<unsigned>(P);
    int* ranks = ws.take<int>(P);
    int* first = ws.take<int>(P);
    int* cnt = ws.take<int>(P);
    int* offsets = ws.take<int>(P);
    int* fill = ws.take<int>(P);
    int* seg = ws.take<int>(P);
    int* counters = ws.take<int>(64);   // [0] n_cells, [1] scratch, [2] kept total
    const bool filter = height_threshold >= 0.f;
    int *keep = nullptr, *newid = nullptr, *zmin = nullptr, *zmax = nullptr;
    const int bw = gx / block_factor, bh = gy / block_factor;
    if (filter) {
        keep = ws.take<int>(max_voxels);
        newid = ws.take<int>(max_voxels);
        zmin = ws.take<int>((size_t)bw * bh);
        zmax = ws.take<int>((size_t)bw * bh);
    }
    if (!cells || !ptbits || !block_sums || !seg || !counters || (filter && !zmax)) {
        set_last_error("rslo_voxelize: workspace too small", cudaErrorMemoryAllocation);
        return (int)cudaErrorMemoryAllocation;
    }
    VoxParams vp{vs[0], vs[1], vs[2], rg[0], rg[1], rg[2], gx, gy, gz};
    const int T = 256, GP = cdiv(P, T);

    RSLO_CHECK(cudaMemsetAsync(cells, 0, nwords * sizeof(uint2), st));
    RSLO_CHECK(cudaMemsetAsync(ptbits, 0, pwords * sizeof(uint2), st));
    RSLO_CHECK(cudaMemsetAsync(first, 0x7f, (size_t)P * sizeof(int), st));
    RSLO_CHECK(cudaMemsetAsync(cnt, 0, (size_t)P * sizeof(int), st));
    RSLO_CHECK(cudaMemsetAsync(fill, 0, (size_t)P * sizeof(int), st));
    RSLO_COUNT();
    k_vox_mark<<<GP, T, 0, st>>>(points, P, F, vp, keys, cells);
    int rc = scan_cells(cells, (int)nwords, block_sums, counters + 0, st);
    if (rc) return rc;
    RSLO_COUNT();
    k_vox_first<<<GP, T, 0, st>>>(keys, P, cells, ranks, first, cnt);
    RSLO_COUNT();
    k_vox_mark_first<<<GP, T, 0, st>>>(first, counters + 0, ptbits);
    rc = scan_cells(ptbits, (int)pwords, block_sums, counters + 1, st);
    if (rc) return rc;
    rc = scan_ints(cnt, offsets, P, block_sums, nullptr, st);
    if (rc) return rc;
    if (filter) {
        RSLO_CHECK(cudaMemsetAsync(zmin, 0x7f, (size_t)bw * bh * sizeof(int), st));   // > any real key
        RSLO_CHECK(cudaMemsetAsync(zmax, 0x80, (size_t)bw * bh * sizeof(int), st));   // < any real key
        RSLO_CHECK(cudaMemsetAsync(keep, 0, (size_t)max_voxels * sizeof(int), st));
    }
    RSLO_COUNT();
    k_vox_scatter<<<GP, T, 0, st>>>(points, P, F, ranks, first, ptbits, offsets, fill, seg, max_voxels,
                                    keys, gx, gy, block_factor, bw, bh, zmin, zmax);
    if (filter) {
        RSLO_COUNT();
        k_vox_filter<<<GP, T, 0, st>>>(counters + 0, first, ptbits, keys, max_voxels, gx, gy,
                                       block_factor, block_size, bw, bh, zmin, zmax, height_threshold,
                                       keep);
        rc = scan_ints(keep, newid, max_voxels, block_sums, counters + 2, st);
        if (rc) return rc;
    }
    RSLO_COUNT();
    if (max_points <= 10)
        k_vox_gather<10><<<cdiv(P, 128), 128, 0, st>>>(points, F, counters + 0, first, ptbits, offsets, cnt, seg,
                                                      keys, max_points, max_voxels, gx, gy, keep, newid,
                                                      batch_idx, voxels, coors, coor_stride, num_points, mean,
                                                      perm);
    else
        k_vox_gather<MAXP_LIMIT><<<cdiv(P, 128), 128, 0, st>>>(points, F, counters + 0, first, ptbits, offsets, cnt,
                                                              seg, keys, max_points, max_voxels, gx, gy, keep,
                                                              newid, batch_idx, voxels, coors, coor_stride,
                                                              num_points, mean, perm);
    RSLO_COUNT();
    k_vox_count<<<1, 32, 0, st>>>(counters + 0, max_voxels, filter ? counters + 2 : nullptr,
                                  n_voxels_dev);
    RSLO_CHECK_LAUNCH("rslo_voxelize");
    return 0;
}

extern "C" int rslo_vfe_mean(const float* voxels, const int32_t* num_points, int n, int max_points,
                             int F, float* mean, rslo_stream_t stream)
{
    if (n <= 0) return 0;
    if (F < 7) {
        set_last_error("rslo_vfe_mean: F < 7", cudaErrorInvalidValue);
        return (int)cudaErrorInvalidValue;
    }
    RSLO_COUNT();
    k_vfe_mean<<<cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(voxels, num_points, n, max_points, F, mean);
    RSLO_CHECK_LAUNCH("rslo_vfe_mean");
    return 0;
}
