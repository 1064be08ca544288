/*
 * oracle.c — CPU restatement of the integer / index arithmetic on RSLO's hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in rslo_b200/ may link, load or call this file; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may.
 *
 * Parity status: the voxeliser and the sparse-conv index generation live in the un-vendored
 * third-party fork DecaYale/spconv_plus (cloned unpinned, reference Dockerfile:58).  They are
 * restated here from the published spconv 1.x algorithm => "parity unpinned" for those rows
 * (SURVEY.md §8c).  The nearest-neighbour search restates code that IS in the reference tree and is
 * pinned against it (tests/test_cpu_oracle.py, oracle/_ref).
 *
 * Plain C99, no dependencies.  Built by oracle/build.py into oracle/_build/liboracle.so.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * a1: spconv.utils.VoxelGenerator.generate, called at rslo/data/preprocess.py:493 through
 * rslo/builder/voxel_builder.py:48-54 (block_filtering=True, voxel_builder.py:75).
 * Published algorithm (spconv 1.x points_to_voxel_3d_with_filtering): one sequential scan over the
 * points in input order.
 *   c_j = floor((p_j - range_min_j) / voxel_size_j)  in float32, point dropped if any c_j is out of
 *   the grid; coordinate stored (z,y,x); a voxel id is handed out at first sight, points of unseen
 *   voxels are skipped once voxel_num >= max_voxels; a point is appended while num < max_points;
 *   per (y,x)-block min/max z is tracked over the points that reach a voxel; post-pass keeps voxel
 *   i iff (max - min over a block_size window centred on its block) > height_threshold.
 * Outputs are compacted by that mask, order preserved.
 * grid: dense int32 scratch of gx*gy*gz entries, all -1 on entry and on exit.
 * Returns the number of kept voxels.
 * ---------------------------------------------------------------------------------------------- */
int oracle_voxelize(const float *points, int P, int F, const float *vsize, const float *range,
                    int gx, int gy, int gz, int max_points, int max_voxels, int block_factor,
                    int block_size, float height_threshold, int32_t *grid, float *voxels,
                    int32_t *coors, int32_t *num_points)
{
    int bh = gy / block_factor, bw = gx / block_factor;
    float *mins = (float *)malloc(sizeof(float) * bh * bw);
    float *maxs = (float *)malloc(sizeof(float) * bh * bw);
    for (int i = 0; i < bh * bw; ++i) { mins[i] = INFINITY; maxs[i] = -INFINITY; }
    int gs[3] = {gx, gy, gz};
    int voxel_num = 0;
    for (int i = 0; i < P; ++i) {
        const float *p = points + (size_t)i * F;
        int c[3];
        int failed = 0;
        for (int j = 0; j < 3; ++j) {
            float q = (p[j] - range[j]) / vsize[j];
            int cj = (int)floorf(q);
            if (cj < 0 || cj >= gs[j]) { failed = 1; break; }
            c[j] = cj;
        }
        if (failed) continue;
        size_t lin = ((size_t)c[2] * gy + c[1]) * gx + c[0];
        int vid = grid[lin];
        if (vid == -1) {
            if (voxel_num >= max_voxels) continue;
            vid = voxel_num++;
            grid[lin] = vid;
            coors[vid * 3 + 0] = c[2];
            coors[vid * 3 + 1] = c[1];
            coors[vid * 3 + 2] = c[0];
            num_points[vid] = 0;
        }
        int by = c[1] / block_factor, bx = c[0] / block_factor;
        if (by < bh && bx < bw) {
            float z = p[2];
            if (z < mins[by * bw + bx]) mins[by * bw + bx] = z;
            if (z > maxs[by * bw + bx]) maxs[by * bw + bx] = z;
        }
        int num = num_points[vid];
        if (num < max_points) {
            memcpy(voxels + ((size_t)vid * max_points + num) * F, p, sizeof(float) * F);
            num_points[vid] = num + 1;
        }
    }
    /* block filter + in-place compaction */
    int kept = 0;
    for (int v = 0; v < voxel_num; ++v) {
        int z = coors[v * 3 + 0], y = coors[v * 3 + 1], x = coors[v * 3 + 2];
        grid[((size_t)z * gy + y) * gx + x] = -1;
        int by = y / block_factor, bx = x / block_factor;
        int y0 = by - block_size / 2, y1 = by + block_size - block_size / 2;
        int x0 = bx - block_size / 2, x1 = bx + block_size - block_size / 2;
        if (y0 < 0) y0 = 0;
        if (x0 < 0) x0 = 0;
        if (y1 > bh) y1 = bh;
        if (x1 > bw) x1 = bw;
        float mx = -INFINITY, mn = INFINITY;
        for (int yy = y0; yy < y1; ++yy)
            for (int xx = x0; xx < x1; ++xx) {
                if (maxs[yy * bw + xx] > mx) mx = maxs[yy * bw + xx];
                if (mins[yy * bw + xx] < mn) mn = mins[yy * bw + xx];
            }
        if ((mx - mn) > height_threshold) {
            if (kept != v) {
                memmove(voxels + (size_t)kept * max_points * F, voxels + (size_t)v * max_points * F,
                        sizeof(float) * max_points * F);
                coors[kept * 3 + 0] = z; coors[kept * 3 + 1] = y; coors[kept * 3 + 2] = x;
                num_points[kept] = num_points[v];
            }
            ++kept;
        }
    }
    free(mins);
    free(maxs);
    return kept;
}

/* ------------------------------------------------------------------------------------------------
 * a10: brute-force nearest neighbour.  Follows ChamferDistanceKernel
 * (thirdparty/chamfer_distance/chamfer_distance.cu:6-137): for every query the argmin over the
 * targets of the squared distance, strict '<' so the lowest index wins ties (also across the
 * kernel's 512-point tiles, :129).  fused != 0 evaluates d the way nvcc contracts the kernel's
 * `x2*x2+y2*y2+z2*z2` for sm_100a (read off the SASS of the reference build,
 * FMUL y2,y2 -> FFMA x2,x2 -> FFMA z2,z2): fma(z2,z2, fma(x2,x2, y2*y2)); fused == 0 evaluates it the way
 * the reference's CPU twin does (chamfer_distance.cpp:116-144, unfused float products and sums).
 * ---------------------------------------------------------------------------------------------- */
void oracle_nn(const float *q, int n, const float *t, int m, int fused, float *dist, int32_t *idx)
{
    for (int j = 0; j < n; ++j) {
        float x1 = q[j * 3 + 0], y1 = q[j * 3 + 1], z1 = q[j * 3 + 2];
        float best = 0.f;
        int besti = 0;
        for (int k = 0; k < m; ++k) {
            float x2 = t[k * 3 + 0] - x1, y2 = t[k * 3 + 1] - y1, z2 = t[k * 3 + 2] - z1;
            float d;
            if (fused) {
                d = fmaf(z2, z2, fmaf(x2, x2, y2 * y2));
            } else {
                volatile float a = x2 * x2, b = y2 * y2, c = z2 * z2;
                volatile float ab = a + b;
                d = ab + c;
            }
            if (k == 0 || d < best) { best = d; besti = k; }
        }
        dist[j] = best;
        idx[j] = besti;
    }
}

/* ------------------------------------------------------------------------------------------------
 * a5: sparse-conv index generation (spconv 1.x ops.get_indice_pairs, un-vendored).
 * Coordinates are (z,y,x) int32, one batch.  The tables are "output-stationary": nbr[o*K + k] is
 * the input row feeding output row o through kernel offset k (row-major kz,ky,kx), or -1.
 *
 * Submanifold (SubMConv3d): outputs = inputs, in input order; in = out + (k - ksize/2).
 * Strided   (SparseConv3d): cross-correlation, in = out*stride - pad + k; the output site set is
 *   the sorted-unique linear index ((z*H)+y)*W+x of every reachable output (spconv 1.x GPU path
 *   ordering); out spatial shape floor((s + 2p - k)/stride) + 1.
 * Inverse   (SparseInverseConv3d): reuses the pairs of the keyed strided conv with roles swapped;
 *   outputs = that conv's inputs; nbr_inv[i*K + k] = o  iff  nbr[o*K + k] = i.
 * ---------------------------------------------------------------------------------------------- */
static int32_t *dense_index(const int32_t *coors, int n, const int shape[3])
{
    size_t tot = (size_t)shape[0] * shape[1] * shape[2];
    int32_t *g = (int32_t *)malloc(sizeof(int32_t) * tot);
    memset(g, 0xff, sizeof(int32_t) * tot);
    for (int i = 0; i < n; ++i)
        g[((size_t)coors[i * 3] * shape[1] + coors[i * 3 + 1]) * shape[2] + coors[i * 3 + 2]] = i;
    return g;
}

void oracle_subm_table(const int32_t *coors, int n, const int *shape, const int *ksize, int32_t *nbr)
{
    int32_t *g = dense_index(coors, n, shape);
    int K = ksize[0] * ksize[1] * ksize[2];
    for (int o = 0; o < n; ++o) {
        int k = 0;
        for (int kz = 0; kz < ksize[0]; ++kz)
            for (int ky = 0; ky < ksize[1]; ++ky)
                for (int kx = 0; kx < ksize[2]; ++kx, ++k) {
                    int z = coors[o * 3] + kz - ksize[0] / 2;
                    int y = coors[o * 3 + 1] + ky - ksize[1] / 2;
                    int x = coors[o * 3 + 2] + kx - ksize[2] / 2;
                    int32_t v = -1;
                    if (z >= 0 && z < shape[0] && y >= 0 && y < shape[1] && x >= 0 && x < shape[2])
                        v = g[((size_t)z * shape[1] + y) * shape[2] + x];
                    nbr[(size_t)o * K + k] = v;
                }
    }
    free(g);
}

/* Returns n_out; out_coors must hold up to n*K rows, nbr up to n*K*K entries (caller trims). */
int oracle_strided_table(const int32_t *coors, int n, const int *shape, const int *ksize,
                         const int *stride, const int *pad, int *out_shape, int32_t *out_coors,
                         int32_t *nbr, int32_t *nbr_inv)
{
    int K = ksize[0] * ksize[1] * ksize[2];
    for (int a = 0; a < 3; ++a) out_shape[a] = (shape[a] + 2 * pad[a] - ksize[a]) / stride[a] + 1;
    size_t tot = (size_t)out_shape[0] * out_shape[1] * out_shape[2];
    int32_t *og = (int32_t *)malloc(sizeof(int32_t) * tot);
    memset(og, 0xff, sizeof(int32_t) * tot);
    /* mark reachable outputs */
    for (int i = 0; i < n; ++i) {
        for (int kz = 0; kz < ksize[0]; ++kz)
            for (int ky = 0; ky < ksize[1]; ++ky)
                for (int kx = 0; kx < ksize[2]; ++kx) {
                    int kk[3] = {kz, ky, kx};
                    int o[3];
                    int ok = 1;
                    for (int a = 0; a < 3; ++a) {
                        int v = coors[i * 3 + a] + pad[a] - kk[a];
                        if (v < 0 || v % stride[a]) { ok = 0; break; }
                        v /= stride[a];
                        if (v >= out_shape[a]) { ok = 0; break; }
                        o[a] = v;
                    }
                    if (ok) og[((size_t)o[0] * out_shape[1] + o[1]) * out_shape[2] + o[2]] = 0;
                }
    }
    /* sorted-unique numbering */
    int n_out = 0;
    for (size_t l = 0; l < tot; ++l)
        if (og[l] == 0) {
            og[l] = n_out;
            out_coors[n_out * 3 + 0] = (int32_t)(l / ((size_t)out_shape[1] * out_shape[2]));
            out_coors[n_out * 3 + 1] = (int32_t)((l / out_shape[2]) % out_shape[1]);
            out_coors[n_out * 3 + 2] = (int32_t)(l % out_shape[2]);
            ++n_out;
        }
    memset(nbr, 0xff, sizeof(int32_t) * (size_t)n_out * K);
    memset(nbr_inv, 0xff, sizeof(int32_t) * (size_t)n * K);
    for (int i = 0; i < n; ++i) {
        int k = 0;
        for (int kz = 0; kz < ksize[0]; ++kz)
            for (int ky = 0; ky < ksize[1]; ++ky)
                for (int kx = 0; kx < ksize[2]; ++kx, ++k) {
                    int kk[3] = {kz, ky, kx};
                    int o[3];
                    int ok = 1;
                    for (int a = 0; a < 3; ++a) {
                        int v = coors[i * 3 + a] + pad[a] - kk[a];
                        if (v < 0 || v % stride[a]) { ok = 0; break; }
                        v /= stride[a];
                        if (v >= out_shape[a]) { ok = 0; break; }
                        o[a] = v;
                    }
                    if (!ok) continue;
                    int32_t oi = og[((size_t)o[0] * out_shape[1] + o[1]) * out_shape[2] + o[2]];
                    nbr[(size_t)oi * K + k] = i;
                    nbr_inv[(size_t)i * K + k] = oi;
                }
    }
    free(og);
    return n_out;
}
