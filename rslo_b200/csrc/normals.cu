// Point-cloud normal estimation on the device (SURVEY §8 row f-N3), sm_100a.
//
// Replaces the offline open3d step of `script/create_hdf5.py:130-147` (called `:322`):
//     pcd.estimate_normals(KDTreeSearchParamHybrid(radius=0.6, max_nn=30)); pcd.orient_normals_towards_camera_location(c)
// open3d 0.9-0.13 semantics restated (the package is not installed in the image: PARITY UNPINNED, see oracle/normals.py):
// neighbours of a point = its <= max_nn nearest points within `radius` (the point itself included); with >= 3 of them
// the normal is the unit eigenvector of the smallest eigenvalue of their covariance (cumulants in double, divided by
// the count), otherwise (0,0,1); then every normal is flipped to face the camera: n <- -n when n . (c - p) < 0.
//
// Method: counting sort of the points into a uniform grid whose cell edge is >= radius (edge and extents are chosen on
// the device from the cloud's bounding box so that the grid fits a fixed cell budget: no host round trip), then one
// thread per point scans its 27 cells and keeps the nearest candidates in a (distance, index)-ordered list, so the
// neighbour set and the summation order do not depend on the order the atomics filled the cells (deterministic).
#include <float.h>

#include "common.cuh"

namespace rslo {
namespace {

constexpr int NRM_MAX_NN = 32;
constexpr int NRM_CELLS = 1 << 22;        // cell budget (16 MB of counters); KITTI at 0.6 m needs ~2.4 M

struct Grid {
    float ox, oy, oz, inv;     // origin, 1 / cell edge
    int nx, ny, nz, pad;
};

__device__ __forceinline__ int fkey(float f)       // order-preserving float -> int
{
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float unkey(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void k_nrm_bbox(const float* __restrict__ p, int ld, int n, int* __restrict__ bb /* min xyz, max xyz */)
{
    int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        for (int k = 0; k < 3; ++k) {
            const int v = fkey(__ldg(p + (size_t)i * ld + k));
            lo[k] = min(lo[k], v);
            hi[k] = max(hi[k], v);
        }
    for (int k = 0; k < 3; ++k) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = min(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = max(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(bb + k, lo[k]);
            atomicMax(bb + 3 + k, hi[k]);
        }
    }
}

__global__ void k_nrm_grid(const int* __restrict__ bb, float radius, Grid* __restrict__ g)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const float lo[3] = {unkey(bb[0]), unkey(bb[1]), unkey(bb[2])};
    const float ex[3] = {unkey(bb[3]) - lo[0], unkey(bb[4]) - lo[1], unkey(bb[5]) - lo[2]};
    float edge = radius;
    for (int it = 0; it < 64; ++it) {                       // grow the cell until the grid fits the budget
        const double cells = (double)((int)(ex[0] / edge) + 1) * ((int)(ex[1] / edge) + 1) * ((int)(ex[2] / edge) + 1);
        if (cells <= (double)NRM_CELLS) break;
        edge *= 1.26f;
    }
    g->ox = lo[0]; g->oy = lo[1]; g->oz = lo[2];
    g->inv = 1.f / edge;
    g->nx = (int)(ex[0] / edge) + 1; g->ny = (int)(ex[1] / edge) + 1; g->nz = (int)(ex[2] / edge) + 1;
    g->pad = 0;
}

__device__ __forceinline__ void cell_of(const Grid& g, float x, float y, float z, int& cx, int& cy, int& cz)
{
    cx = min(max((int)((x - g.ox) * g.inv), 0), g.nx - 1);
    cy = min(max((int)((y - g.oy) * g.inv), 0), g.ny - 1);
    cz = min(max((int)((z - g.oz) * g.inv), 0), g.nz - 1);
}

__global__ void k_nrm_count(const float* __restrict__ p, int ld, int n, const Grid* __restrict__ gp, int* __restrict__ cnt)
{
    const Grid g = *gp;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int cx, cy, cz;
        cell_of(g, __ldg(p + (size_t)i * ld), __ldg(p + (size_t)i * ld + 1), __ldg(p + (size_t)i * ld + 2), cx, cy, cz);
        atomicAdd(cnt + ((size_t)cz * g.ny + cy) * g.nx + cx, 1);
    }
}

__global__ void k_nrm_fill(const float* __restrict__ p, int ld, int n, const Grid* __restrict__ gp, const int* __restrict__ start,
                           int* __restrict__ cursor, int* __restrict__ order)
{
    const Grid g = *gp;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int cx, cy, cz;
        cell_of(g, __ldg(p + (size_t)i * ld), __ldg(p + (size_t)i * ld + 1), __ldg(p + (size_t)i * ld + 2), cx, cy, cz);
        const size_t c = ((size_t)cz * g.ny + cy) * g.nx + cx;
        order[start[c] + atomicAdd(cursor + c, 1)] = i;
    }
}

// unit eigenvector of the smallest eigenvalue of the symmetric matrix {a00,a01,a02,a11,a12,a22}: cyclic Jacobi in double
__device__ void smallest_eigenvector(double a00, double a01, double a02, double a11, double a12, double a22, double* v)
{
    double A[3][3] = {{a00, a01, a02}, {a01, a11, a12}, {a02, a12, a22}};
    double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 12; ++sweep) {
        const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        if (off < 1e-300) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (fabs(A[p][q]) < 1e-300) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq;
                    A[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - s * aqk;
                    A[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq;
                    V[k][q] = s * vkp + c * vkq;
                }
            }
    }
    int m = 0;
    if (A[1][1] < A[m][m]) m = 1;
    if (A[2][2] < A[m][m]) m = 2;
    const double nn = sqrt(V[0][m] * V[0][m] + V[1][m] * V[1][m] + V[2][m] * V[2][m]);
    v[0] = V[0][m] / nn; v[1] = V[1][m] / nn; v[2] = V[2][m] / nn;
}

__global__ void __launch_bounds__(128)
k_nrm_estimate(const float* __restrict__ p, int ld, int n, const Grid* __restrict__ gp, const int* __restrict__ start,
               const int* __restrict__ cnt, const int* __restrict__ order, float radius, int max_nn, float camx, float camy,
               float camz, float* __restrict__ normals)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Grid g = *gp;
    const float x = __ldg(p + (size_t)i * ld), y = __ldg(p + (size_t)i * ld + 1), z = __ldg(p + (size_t)i * ld + 2);
    int cx, cy, cz;
    cell_of(g, x, y, z, cx, cy, cz);
    const float r2 = radius * radius;
    float bd[NRM_MAX_NN];
    int bi[NRM_MAX_NN];
    int m = 0;
    for (int dz = -1; dz <= 1; ++dz) {
        const int zz = cz + dz;
        if (zz < 0 || zz >= g.nz) continue;
        for (int dy = -1; dy <= 1; ++dy) {
            const int yy = cy + dy;
            if (yy < 0 || yy >= g.ny) continue;
            const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.nx - 1);
            const size_t row = ((size_t)zz * g.ny + yy) * g.nx;
            const int s0 = start[row + x0], s1 = start[row + x1] + cnt[row + x1];      // the three cells are contiguous
            for (int s = s0; s < s1; ++s) {
                const int j = order[s];
                const float ax = __ldg(p + (size_t)j * ld) - x, ay = __ldg(p + (size_t)j * ld + 1) - y,
                            az = __ldg(p + (size_t)j * ld + 2) - z;
                const float d = (ax * ax + ay * ay) + az * az;
                if (d > r2) continue;
                if (m == max_nn && !(d < bd[m - 1] || (d == bd[m - 1] && j < bi[m - 1]))) continue;
                int k = m < max_nn ? m++ : max_nn - 1;              // insertion into the (distance, index)-ordered list
                while (k > 0 && (bd[k - 1] > d || (bd[k - 1] == d && bi[k - 1] > j))) {
                    bd[k] = bd[k - 1];
                    bi[k] = bi[k - 1];
                    --k;
                }
                bd[k] = d;
                bi[k] = j;
            }
        }
    }
    double nv[3] = {0.0, 0.0, 1.0};
    if (m >= 3) {
        double c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int k = 0; k < m; ++k) {
            const int j = bi[k];
            const double px = __ldg(p + (size_t)j * ld), py = __ldg(p + (size_t)j * ld + 1), pz = __ldg(p + (size_t)j * ld + 2);
            c[0] += px; c[1] += py; c[2] += pz;
            c[3] += px * px; c[4] += px * py; c[5] += px * pz; c[6] += py * py; c[7] += py * pz; c[8] += pz * pz;
        }
        for (int k = 0; k < 9; ++k) c[k] /= (double)m;
        smallest_eigenvector(c[3] - c[0] * c[0], c[4] - c[0] * c[1], c[5] - c[0] * c[2], c[6] - c[1] * c[1], c[7] - c[1] * c[2],
                             c[8] - c[2] * c[2], nv);
    }
    if (nv[0] * ((double)camx - x) + nv[1] * ((double)camy - y) + nv[2] * ((double)camz - z) < 0.0) {
        nv[0] = -nv[0]; nv[1] = -nv[1]; nv[2] = -nv[2];
    }
    normals[(size_t)i * 3] = (float)nv[0];
    normals[(size_t)i * 3 + 1] = (float)nv[1];
    normals[(size_t)i * 3 + 2] = (float)nv[2];
}

}  // namespace
}  // namespace rslo

using namespace rslo;

extern "C" size_t rslo_estimate_normals_workspace_bytes(int n)
{
    return ws_round(8 * sizeof(int)) + ws_round(sizeof(Grid)) + 3 * ws_round((size_t)NRM_CELLS * sizeof(int)) +
           ws_round((size_t)(n > 0 ? n : 1) * sizeof(int)) + ws_round(scan_ws_ints(NRM_CELLS) * sizeof(int)) + 1024;
}

extern "C" int rslo_estimate_normals(const float* xyz, int ld, int n, float radius, int max_nn, const float* camera_host,
                                     float* normals, void* workspace, size_t workspace_bytes, rslo_stream_t stream)
{
    if (n <= 0) return 0;
    if (max_nn < 3 || max_nn > NRM_MAX_NN || radius <= 0.f || ld < 3 || workspace == nullptr ||
        workspace_bytes < rslo_estimate_normals_workspace_bytes(n)) {
        set_last_error("rslo_estimate_normals: bad arguments (3 <= max_nn <= 32) or workspace too small", cudaErrorInvalidValue);
        return (int)cudaErrorInvalidValue;
    }
    cudaStream_t st = (cudaStream_t)stream;
    Workspace ws(workspace, workspace_bytes);
    int* bb = ws.take<int>(8);
    Grid* grid = ws.take<Grid>(1);
    int* cnt = ws.take<int>(NRM_CELLS);
    int* start = ws.take<int>(NRM_CELLS);
    int* cursor = ws.take<int>(NRM_CELLS);
    int* order = ws.take<int>(n);
    int* block_sums = ws.take<int>(scan_ws_ints(NRM_CELLS));
    RSLO_CHECK(cudaMemsetAsync(bb, 0x7f, 3 * sizeof(int), st));            // > any key
    RSLO_CHECK(cudaMemsetAsync(bb + 3, 0x80, 3 * sizeof(int), st));        // < any key
    RSLO_CHECK(cudaMemsetAsync(cnt, 0, (size_t)NRM_CELLS * sizeof(int), st));
    RSLO_CHECK(cudaMemsetAsync(cursor, 0, (size_t)NRM_CELLS * sizeof(int), st));
    int blocks = cdiv(n, 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    RSLO_COUNT();
    k_nrm_bbox<<<blocks, 256, 0, st>>>(xyz, ld, n, bb);
    RSLO_COUNT();
    k_nrm_grid<<<1, 32, 0, st>>>(bb, radius, grid);
    RSLO_COUNT();
    k_nrm_count<<<blocks, 256, 0, st>>>(xyz, ld, n, grid, cnt);
    int rc = scan_ints(cnt, start, NRM_CELLS, block_sums, nullptr, st);
    if (rc) return rc;
    RSLO_COUNT();
    k_nrm_fill<<<blocks, 256, 0, st>>>(xyz, ld, n, grid, start, cursor, order);
    RSLO_COUNT();
    k_nrm_estimate<<<cdiv(n, 128), 128, 0, st>>>(xyz, ld, n, grid, start, cnt, order, radius, max_nn, camera_host[0],
                                                camera_host[1], camera_host[2], normals);
    RSLO_CHECK_LAUNCH("rslo_estimate_normals");
    return 0;
}
