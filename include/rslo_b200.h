/*
 * rslo_b200.h — C ABI of the B200-native kernels behind RSLO's per-frame-pair hot path.
 *
 * The reference has exactly one native boundary on this path, the pybind module `cd`
 * (thirdparty/chamfer_distance/chamfer_distance.cpp:237-244): caller-allocated contiguous
 * float32 / int32 buffers, written in place, void return, errors printf'd, legacy default stream.
 * Everything else the reference reaches through Python calls into the un-vendored spconv fork
 * (rslo/builder/voxel_builder.py:48-54, rslo/models/middle.py:119-245).  This header keeps the
 * `cd` calling convention (borrowed device pointers, no allocation inside, in-place outputs) and
 * extends it to the other operators of the path, with three changes: every entry point takes the
 * CUDA stream to run on, returns 0 or a cudaError_t value (text via rslo_last_error()), and scratch
 * memory comes from a caller-provided workspace sized by the matching *_workspace_bytes().
 *
 * All pointers are DEVICE pointers unless the name ends in _host.  Row counts that depend on the
 * data are passed twice: `n_cap` (rows allocated / grid size) and `n_dev` (device int holding the
 * live count, or NULL meaning n_cap), so a frame can be enqueued without host round trips.
 * No torch types appear here; the Python side (rslo_b200/_lib.py) binds this with ctypes.
 */
#ifndef RSLO_B200_H
#define RSLO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* rslo_stream_t; /* cudaStream_t */

int rslo_abi_version(void);
const char* rslo_last_error(void);
/* Kernels launched by this library in the calling process so far (measurement aid). */
unsigned long long rslo_kernel_launch_count(void);
/* The caller is about to capture (on != 0) / has finished capturing (0) the head's launches into a CUDA graph: while
 * set, the head's kernels are launched with the programmatic-dependent-launch attribute (each of them waits for its
 * predecessor with griddepcontrol.wait before touching global memory), which removes the launch latency between the
 * ~880 kernel nodes of the two graphs.  Off by default: on an eager path the attribute costs host time per launch. */
void rslo_set_graph_capture_hint(int on);

/* ---- a10: exact nearest neighbour ---------------------------------------------------------------
 * Replaces cd.forward_cuda_one_direction (chamfer_distance.cpp:237-244 ->
 * ChamferDistanceKernel, chamfer_distance.cu:6-137) for batch 1: for each of n query points the
 * squared distance to, and index of, its nearest of m target points; bit-identical to the reference
 * kernel's output (d = fma(z,z,fma(x,x,y*y)), lowest index wins ties).  Uniform-grid search with a
 * conservative termination bound; queries that cannot be bounded fall back to an exhaustive scan. */
size_t rslo_nn_workspace_bytes(int n, int m);
int rslo_nn_exact(const float* query, int n, const float* target, int m, float* dist, int32_t* idx,
                  void* workspace, size_t workspace_bytes, rslo_stream_t stream);
/* Brute-force tiled variant (same results); kept as the in-library cross-check and for tiny m. */
int rslo_nn_brute(const float* query, int n, const float* target, int m, float* dist, int32_t* idx,
                  rslo_stream_t stream);

/* ---- a1 + a3: voxeliser (+ fused VFE mean) ------------------------------------------------------
 * Replaces spconv.utils.VoxelGenerator.generate as called through _VoxelGenerator.generate
 * (rslo/builder/voxel_builder.py:48-54; call site rslo/data/preprocess.py:493) and, when `mean` is
 * given, SimpleVoxel_XYZINormalC.forward (rslo/models/voxel_encoder.py:272-280).
 * points [P,F] f32.  Outputs (rows in first-come order, exactly the sequential scan's):
 *   voxels   [max_voxels,max_points,F] f32 zero padded, or NULL to skip materialising it;
 *   coors    [max_voxels,coor_stride] i32, last three columns (z,y,x); with coor_stride 4 column 0
 *            is set to batch_idx (merge_second_batch's pad, rslo/data/preprocess.py:84-85);
 *   num_points [max_voxels] i32;  mean [max_voxels,7] f32 or NULL;
 *   n_voxels_dev: device int, number of voxels produced;
 *   cells / perm: optional (may be NULL) site table of the produced voxels over a
 *            (table_d, gy, gx) cell grid for rslo_subm_table / rslo_strided_sites — cells is
 *            2*ceil(table_d*gy*gx/32) uint32, perm is P int32.
 * height_threshold < 0 keeps every voxel (the shipped configs: -1). */
size_t rslo_voxelize_workspace_bytes(int P, int gx, int gy, int table_d, int max_voxels);
int rslo_voxelize(const float* points, int P, int F, const float* voxel_size_host,
                  const float* pc_range_host, int gx, int gy, int gz, int max_points, int max_voxels,
                  int block_factor, int block_size, float height_threshold, int batch_idx,
                  float* voxels, int32_t* coors, int coor_stride, int32_t* num_points, float* mean,
                  int32_t* n_voxels_dev, uint32_t* cells, int32_t* perm, int table_d,
                  void* workspace, size_t workspace_bytes, rslo_stream_t stream);

/* VFE alone on materialised voxels (voxel_encoder.py:272-280). */
int rslo_vfe_mean(const float* voxels, const int32_t* num_points, int n, int max_points, int F,
                  float* mean, rslo_stream_t stream);

/* ---- a5: sparse-conv index generation -----------------------------------------------------------
 * Replaces spconv's get_indice_pairs behind SubMConv3d / SparseConv3d / SparseInverseConv3d
 * (rslo/models/middle.py:119-213).  Tables are output-stationary: nbr[o*K+k] = input row or -1,
 * K = kd*kh*kw, k row-major (kz,ky,kx).
 * Site table of a level: cells = 2*ceil(D*H*W/32) uint32 ({bitmap word, rank of its first bit}),
 * perm = rank -> row (NULL when rows are already in sorted cell order). */
size_t rslo_site_table_workspace_bytes(int D, int H, int W);
int rslo_site_table_build(const int32_t* coors, int coor_stride, int n_cap, const int32_t* n_dev,
                          int D, int H, int W, uint32_t* cells, int32_t* perm, void* workspace,
                          size_t workspace_bytes, rslo_stream_t stream);
int rslo_subm_table(const int32_t* coors, int coor_stride, int n_cap, const int32_t* n_dev, int D,
                    int H, int W, const uint32_t* cells, const int32_t* perm, int kd, int kh, int kw,
                    int32_t* nbr, rslo_stream_t stream);
/* Strided conv: output site set (sorted by cell index), its site table, and both tables.
 * out_coors [out_cap,4] (b,z,y,x); nbr [out_cap,K]; nbr_inv [n_cap,K] (the SparseInverseConv3d
 * table: nbr_inv[i*K+k] = o iff nbr[o*K+k] = i); n_out_dev: device int[2] = {min(count, out_cap),
 * count} — count > out_cap means the caller must retry with a larger out_cap. */
size_t rslo_strided_workspace_bytes(int oD, int oH, int oW);
int rslo_strided_table(const int32_t* coors, int coor_stride, int n_cap, const int32_t* n_dev,
                       int D, int H, int W, int kd, int kh, int kw, int sd, int sh, int sw, int pd,
                       int ph, int pw, uint32_t* out_cells, int32_t* out_coors, int out_cap,
                       int32_t* n_out_dev, int32_t* nbr, int32_t* nbr_inv, void* workspace,
                       size_t workspace_bytes, rslo_stream_t stream);

/* Several frames share one pass through the encoder: their tables are appended row-wise, row indices of
 * frame f shifted by the number of rows before it.  dst[i] = src[i] >= 0 ? src[i] + add : -1. */
int rslo_table_concat(const int32_t* src, long long count, int add, int32_t* dst, rslo_stream_t stream);
/* the same for nseg (source, destination) segments in one launch per RSLO_CONCAT_MAX segments (host array) */
#define RSLO_CONCAT_MAX 64
typedef struct {
    const int32_t* src;
    int32_t* dst;
    long long count;
    int add;
    int reserved;
} rslo_concat_seg_t;
int rslo_table_concat_multi(const rslo_concat_seg_t* segs_host, int nseg, rslo_stream_t stream);

/* ---- a6: sparse convolution ---------------------------------------------------------------------
 * out[o,:] = act( scale * (bias + sum_k in[nbr[o,k],:] @ W[k]) + shift )
 * Replaces Fsp.indice_conv / indice_inverse_conv + the LeakyReLU (and eval-mode BatchNorm1d of the
 * covariance decoder) that follow each layer in middle.py:119-213.  W [K,Cin,Cout] f32 (the
 * reference's [kD,kH,kW,Cin,Cout] state_dict layout), bias [Cout] or NULL, scale/shift [Cout] or
 * NULL, act: 0 none, 1 LeakyReLU(slope). */
int rslo_spconv_forward(const float* in, const int32_t* nbr, int n_out_cap, const int32_t* n_out_dev,
                        int K, int Cin, int Cout, const float* weight, const float* bias,
                        const float* scale, const float* shift, int act, float slope, float* out,
                        rslo_stream_t stream);
/* Wt[k] = W[k']^T ([K,Cout,Cin]), k' = K-1-k when mirror != 0 else k: the filter bank the data
 * gradient convolves with. */
int rslo_spconv_transpose_weight(const float* weight, int K, int Cin, int Cout, int mirror,
                                 float* weight_t, rslo_stream_t stream);
/* d(in)[i,:] = sum_k g[nbr_t[i,k],:] @ Wt[k] — the data gradient is itself a gather-convolution over
 * the transposed table: nbr_inv for a strided conv, nbr for an inverse conv, and for a submanifold
 * conv the same table with mirrored offsets (Wt built with mirror=1).  Cin/Cout are the FORWARD
 * layer's; grad_out [n_out,Cout], grad_in [n_in,Cin]. */
int rslo_spconv_backward_data(const float* grad_out, const int32_t* nbr_t, int n_in_cap,
                              const int32_t* n_in_dev, int K, int Cin, int Cout, const float* weight_t,
                              float* grad_in, rslo_stream_t stream);
/* dW[k] = sum_o in[nbr[o,k],:]^T (x) g[o,:], dbias[c] = sum_o g[o,c] (grad_bias may be NULL); both are
 * zeroed inside. */
int rslo_spconv_backward_weight(const float* in, const float* grad_out, const int32_t* nbr,
                                int n_out_cap, const int32_t* n_out_dev, int K, int Cin, int Cout,
                                float* grad_weight, float* grad_bias, rslo_stream_t stream);

/* Activation backward fused with the bias gradient: grad_act = grad_out * (out > 0 ? 1 : slope) for
 * act 1 (LeakyReLU, from the saved layer OUTPUT), grad_bias[c] = sum_o grad_act[o,c] (NULL to skip; zeroed
 * inside).  act 0: bias gradient only (out / grad_act may be NULL). */
int rslo_act_backward(const float* grad_out, const float* out, int n_cap, const int32_t* n_dev, int C, int act,
                      float slope, float* grad_act, float* grad_bias, rslo_stream_t stream);

/* Tensor-core variant for Cin, Cout in {32, 64} (csrc/spconv_tc.cu): tcgen05.mma kind::tf32 with a
 * 3xTF32 operand split (FP32-level accuracy), TMEM accumulator, weight tiles staged by bulk async copy.
 * `image` is the pre-swizzled {hi, lo} weight image made by rslo_spconv_tc_prepare
 * (rslo_spconv_tc_image_bytes bytes): transpose = 0 for the forward filter bank (kdim = Cin, ndim = Cout),
 * transpose = 1 (+ mirror for submanifold tables) for the data gradient (kdim = Cout, ndim = Cin).
 * rslo_spconv_tc_forward computes out[o,:] = act(bias + sum_k in[nbr[o,k],:] @ B_k), in [.,kdim], out [.,ndim]. */
int rslo_spconv_tc_supported(int Cin, int Cout, int K);
size_t rslo_spconv_tc_image_bytes(int K, int Cin, int Cout);
int rslo_spconv_tc_prepare(const float* weight, int K, int Cin, int Cout, int transpose, int mirror,
                           float* image, rslo_stream_t stream);
size_t rslo_spconv_tc_workspace_bytes(int n_out_cap, int ndim);
int rslo_spconv_tc_forward(const float* in, const int32_t* nbr, int n_out_cap, const int32_t* n_out_dev,
                           int K, int kdim, int ndim, const float* image, const float* bias, int act,
                           float slope, float* out, void* workspace, size_t workspace_bytes,
                           rslo_stream_t stream);

/* Weight gradient on the tensor cores (csrc/spconv_tc_wgrad.cu): dW[k] = sum_o in[nbr[o,k],:]^T (x) g[o,:]
 * as MN-major tcgen05 MMAs with the offsets stacked along M; grad_weight [K,Cin,Cout] is zeroed inside. */
int rslo_spconv_tc_wgrad_supported(int Cin, int Cout);
int rslo_spconv_tc_backward_weight(const float* in, const float* grad_out, const int32_t* nbr, int n_out_cap,
                                   const int32_t* n_out_dev, int K, int Cin, int Cout, float* grad_weight,
                                   rslo_stream_t stream);

/* ---- a8: dense 2-D convolutions of the odometry head (csrc/conv2d_tc.cu) ---------------------------
 * Replaces cuDNN under rslo/models/odom_pred_base.py:155-276, rslo/layers/MaskConv.py:53-63,
 * rslo/models/custom_resnet_spc.py:224-298: 3x3 (pad 1) and 1x1 (pad 0), stride 1 or 2, groups 1,
 * Cin and Cout multiples of 32.  tcgen05.mma kind::tf32 with the 3xTF32 operand split (FP32-level
 * accuracy), operands staged by TMA tensor tiles (image border = TMA zero fill).
 * Activations are NHWC "split pairs" [2][B][H][W][C]: plane 0 = hi = RN_tf32(x), plane 1 = x - hi
 * (rslo_conv2d_split makes one from a plain NHWC tensor).  Weights are OIHW as in the reference's
 * state_dict; rslo_conv2d_tc_prepare makes the split image [2][k*k][N][Kd] once per weight update
 * (mode 0: forward, N = Cout, Kd = Cin; mode 1: data gradient, N = Cin, Kd = Cout), 8*k*k*Cin*Cout bytes.
 * workspace: rslo_conv2d_tc_workspace_bytes bytes, ZEROED once by the caller before first use and then
 * reused on one stream (split-K tile counters; the kernels leave them zeroed). */
int rslo_conv2d_tc_supported(int Cin, int Cout, int ksize, int stride);
int rslo_conv2d_split(const float* x, size_t n, float* split_pair, rslo_stream_t stream);
int rslo_conv2d_tc_prepare(const float* weight_oihw, int Cout, int Cout_padded, int Cin, int ksize, int mode,
                           float* image, rslo_stream_t stream);
size_t rslo_conv2d_tc_workspace_bytes(int B, int H, int W, int Cmax);
/* y [B][Ho][Wo][Cout] = conv(x) (+ bias, may be NULL) (ReLU when relu != 0).  stats (may be NULL):
 * double [B / imgs_per_group][Cout][2], pre-zeroed; the epilogue adds the per-channel sum and sum of squares of
 * y over each statistics group of images (the BatchNorm batch statistics of the layer that follows).
 * Narrow heads (7 / 1 output channels): prepare with Cout_padded = 32 and pass Cout = 32 here. */
int rslo_conv2d_tc_forward(const float* x_split, int B, int H, int W, int Cin, const float* image, int Cout,
                           int ksize, int stride, const float* bias, int relu, float* y, double* stats,
                           int imgs_per_group, void* workspace, size_t workspace_bytes, rslo_stream_t stream);
/* dx [B][H][W][Cin] (= , or += when accumulate) from g_split [2][B][Ho][Wo][Cout] and the mode-1 image */
int rslo_conv2d_tc_backward_data(const float* g_split, int B, int H, int W, int Cin, const float* image_t, int Cout,
                                 int ksize, int stride, float* dx, int accumulate, void* workspace,
                                 size_t workspace_bytes, rslo_stream_t stream);
/* grad_weight_oihw [Cout_real][Cin][k][k] (= , or += when accumulate); g_split has Cout (padded) channels;
 * scratch: rslo_conv2d_tc_wgrad_scratch_bytes(Cin, Cout, ksize) floats laid out [k*k][Cin][Cout].
 * grad_weight_oihw == NULL: the products are ADDED into scratch (which the caller zeroed) and left there for
 * rslo_conv2d_multi_wgrad_finish — one memset and one finish launch for all convolutions of a backward pass. */
size_t rslo_conv2d_tc_wgrad_scratch_bytes(int Cin, int Cout, int ksize);
int rslo_conv2d_tc_backward_weight(const float* x_split, const float* g_split, int B, int H, int W, int Cin, int Cout,
                                   int ksize, int stride, int Cout_real, float* scratch, int accumulate,
                                   float* grad_weight_oihw, rslo_stream_t stream);

/* Table-driven variants (device arrays of these structs): weight images (+ zero-padded bias) of n convolutions in
 * one launch; [k*k][Cin][CoutP] weight-gradient scratch -> OIHW gradients of n convolutions in one launch. */
typedef struct {
    const float* w;        /* OIHW [Cout][Cin][k][k] */
    float* img_fwd;        /* mode-0 image [2][taps][CoutP][Cin] or NULL */
    float* img_bwd;        /* mode-1 image [2][taps][Cin][CoutP] or NULL */
    const float* bias;     /* [Cout] or NULL */
    float* bias_pad;       /* [CoutP] or NULL */
    int Cout, CoutP, Cin, taps;
} rslo_conv_prep_t;
typedef struct {
    const float* dW;       /* [taps][Cin][CoutP] */
    float* gw;             /* OIHW [Cout][Cin][taps] */
    int Cout, CoutP, Cin, taps;
} rslo_wgrad_finish_t;
int rslo_conv2d_multi_prepare(const rslo_conv_prep_t* table_dev, int n, rslo_stream_t stream);
int rslo_conv2d_multi_wgrad_finish(const rslo_wgrad_finish_t* table_dev, int n, rslo_stream_t stream);

/* ---- a8: what the head does between its convolutions (csrc/head_ops.cu), NHWC -------------------------------
 * Frame-pair input (rslo/models/odom_pred.py:165-170): x1, x2 [B][C][HW] (NCHW BEV maps of the two frames) ->
 * split pair [2][B][HW][2C] of cat(x1, x2) and mask [B][HW] = (sum_c x1 != 0); and the gradient's way back. */
int rslo_head_pack_input(const float* x1, const float* x2, int B, int C, int HW, float* split_pair, float* mask,
                         rslo_stream_t stream);
int rslo_head_unpack_grad(const float* dx, int B, int C, int HW, float* g1, float* g2, rslo_stream_t stream);
/* BatchNorm2d (+ residual) (+ ReLU) over y [B][HW][C] (SPC_SyncBN2d / SPC_ReLU / BasicBlock,
 * rslo/layers/SparseConv.py:96-132, rslo/models/custom_resnet_spc.py:224-298).
 * stats != NULL: batch statistics, double [G][C][2] = {sum, sum of squares} per statistics group of
 * imgs_per_group images (G = B / imgs_per_group) as accumulated by rslo_conv2d_tc_forward; running_mean/var
 * (may be NULL) are updated group after group, update_repeat times each (momentum, unbiased variance), and
 * num_batches_tracked += G * update_repeat.  stats == NULL: running statistics (eval / frozen BN).
 * Outputs (each may be NULL): z [B][HW][C]; z_split [2][B][HW][C]; mean_rstd float [G][C][2] for the backward. */
int rslo_bn_act_forward(const float* y, int B, int HW, int C, int imgs_per_group, const double* stats,
                        const float* gamma, const float* beta, float* running_mean, float* running_var,
                        long long* num_batches_tracked, float eps, float momentum, int update_repeat,
                        const float* residual, int relu, float* z, float* z_split, float* mean_rstd,
                        rslo_stream_t stream);
/* Backward of the above: dz (gradient w.r.t. z), z (only read when relu), y, mean_rstd, gamma ->
 * g_split [2][B][HW][C] (gradient w.r.t. y as a split pair, the operand of the convolution backward kernels),
 * dres (may be NULL; = or += the gradient w.r.t. the residual), dgamma / dbeta [C], dbias [C] (may be NULL: gradient
 * of a bias added before the normalisation).  sums: double [G][C][2] scratch, ZEROED by the caller. */
int rslo_bn_act_backward(const float* dz, const float* z, const float* y, int B, int HW, int C, int imgs_per_group,
                         const float* mean_rstd, const float* gamma, int relu, int batch_stats, double* sums,
                         float* g_split, float* dres, int dres_accumulate, float* dgamma, float* dbeta, float* dbias,
                         rslo_stream_t stream);
/* nn.Upsample(scale_factor=up, nearest) of z [B][H][W][C] written as channels [choff, choff+C) of the split pair
 * [2][B][up*H][up*W][ld] (torch.cat + Upsample of the decoder, rslo/models/odom_pred_base.py:196-207), and back:
 * dz [B][H][W][C] (= or +=) sum over each up x up block of dcat [B][up*H][up*W][ld] channels [choff, choff+C). */
int rslo_upcat_split(const float* z, int B, int H, int W, int C, int up, int ld, int choff, float* dst_split,
                     rslo_stream_t stream);
int rslo_upcat_backward(const float* dcat, int B, int H, int W, int C, int up, int ld, int choff, float* dz,
                        int accumulate, rslo_stream_t stream);
/* out[c] += sum_rows g[row][c] for c < C <= 32, g [N][ld] (bias gradient of the 7- / 1-channel output convs) */
int rslo_bias_grad(const float* g, int N, int ld, int C, float* out, rslo_stream_t stream);

/* ---- a9 / a13: the two tails of the head (csrc/pose_tail.cu), one launch forward + one backward each --------
 * Geometry of the (t,q) map: cell (i,j) has its anchor at x = (j - ox) vsx, y = (-i + oy) vsy, z = (0 - oz) vsz
 * (rslo/data/dataset.py:145-146,169-171). */
typedef struct {
    int H, W;
    float ox, oy, oz, vsx, vsy, vsz;
} rslo_tq_geom_t;
/* Head tail (rslo/models/odom_pred.py:226-313, rslo/layers/confidence.py:23-34, rslo/data/dataset.py:121-208):
 * inputs NHWC with 32-channel rows as written by the trunk's output convolutions: tq32 [B][H][W][32] (7 used),
 * t_logit32 / r_logit32 (1 used), py0_32 [B][H/4][W/4][32], py1_32 [B][H/2][W/2][32] (7 used), mask [B][H][W].
 * outputs (NCHW as in the reference): pose_t [B,3], pose_q [B,4] (w,x,y,z, unit), tq_map_g [B,7,H,W] (masked),
 * t_conf / r_conf [B,1,H,W], pyramid_motion levels 2 (full), 1 (H/2), 0 (H/4): pred [B,7,h,w], mask [B,2,h,w];
 * occ1 [B,H/2,W/2], occ0 [B,H/4,W/4] (max-pooled occupancy) and save [B][16] are kept for the backward. */
int rslo_head_tail_forward(const float* tq32, const float* t_logit32, const float* r_logit32, const float* mask,
                           const float* py0_32, const float* py1_32, int B, rslo_tq_geom_t geom, float* pose_t,
                           float* pose_q, float* tq_map_g, float* t_conf, float* r_conf, float* pm2_pred,
                           float* pm2_mask, float* pm1_pred, float* pm1_mask, float* pm0_pred, float* pm0_mask,
                           float* occ1, float* occ0, float* save, rslo_stream_t stream);
/* gradient inputs may be NULL (no gradient); outputs are full 32-channel rows (zeros in the unused channels) */
int rslo_head_tail_backward(const float* tq32, const float* mask, const float* t_conf, const float* r_conf,
                            const float* occ1, const float* occ0, const float* save, int B, rslo_tq_geom_t geom,
                            const float* g_pose_t, const float* g_pose_q, const float* g_tq_map_g,
                            const float* g_t_conf, const float* g_r_conf, const float* g_pm2_pred,
                            const float* g_pm1_pred, const float* g_pm0_pred, float* d_tq32, float* d_t_logit32,
                            float* d_r_logit32, float* d_py1_32, float* d_py0_32, rslo_stream_t stream);
/* Loss tail (rslo/models/voxel_odom_net.py:727-795, rslo/data/dataset.py:52-116, rslo/core/losses.py:155-197 with
 * focal_gamma 0): pseudo labels R* = res_r R(q_pred) (identity while identity_pose), t* = res_r T + res_t ->
 * target map tq_target [B,7,H,W]; losses8 = {T, R, pyramid T x3 (levels 0,1,2), pyramid R x3}, each
 * w (sum_b e^-alpha L_b / (B + 1e-12) + alpha).  alpha_*: device scalars (may alias).  save: [B][24] floats;
 * counter: one zeroed device int (left zeroed). */
int rslo_loss_tail_forward(const float* T_pred, const float* q_pred, const float* pm0_pred, const float* pm0_mask,
                           const float* pm1_pred, const float* pm1_mask, const float* pm2_pred, const float* pm2_mask,
                           const float* res_r, const float* res_t, int identity_pose, int B, rslo_tq_geom_t geom,
                           const float* alpha_t, const float* alpha_r, const float* alpha_pt, const float* alpha_pr,
                           float w_t, float w_r, float w_pt, float w_pr, float* tq_target, float* save, float* losses8,
                           int* counter, rslo_stream_t stream);
int rslo_loss_tail_backward(const float* T_pred, const float* q_pred, const float* pm0_pred, const float* pm0_mask,
                            const float* pm1_pred, const float* pm1_mask, const float* pm2_pred, const float* pm2_mask,
                            int B, rslo_tq_geom_t geom, const float* alpha_t, const float* alpha_r,
                            const float* alpha_pt, const float* alpha_pr, float w_t, float w_r, float w_pt, float w_pr,
                            const float* save, const float* g_losses8, float* dT, float* dq, float* d_pm0, float* d_pm1,
                            float* d_pm2, float* dalpha4, rslo_stream_t stream);

/* ---- a7: SparseConvTensor.dense() + view (middle.py:240-243) ------------------------------------
 * feat [n,C] at sites of a (D,H,W) level -> dense [C*D, H, W] f32 (zero where no site). */
int rslo_dense_from_sites(const float* feat, int C, const uint32_t* cells, const int32_t* perm,
                          int D, int H, int W, float* dense, rslo_stream_t stream);
/* gradient of the above: grad_feat[r,c] = grad_dense[c, cell(r)] */
int rslo_dense_backward(const float* grad_dense, int C, const int32_t* coors, int coor_stride,
                        int n_cap, const int32_t* n_dev, int D, int H, int W, float* grad_feat,
                        rslo_stream_t stream);

/* ---- a12: weighted Kabsch alignment (no host round trip) -----------------------------------------
 * Replaces SVDHead.forward (rslo/layers/svd.py:13-64) as driven by the ICP refinement in
 * rslo/core/losses.py:440-488.  src [n,3], tgt rows f32 row-major; tgt_idx [n] or NULL: row i pairs with
 * tgt[tgt_idx[i]] (the association gather); weight [n] or NULL (ones); normal [n,3] or NULL: when given the
 * weight is multiplied by cos^2(normal_i, tgt_i - src_i) (losses.py:411); mask [n]
 * 0/1 floats or NULL; optional ROI test dist[i] < *dist_threshold (both device, may be NULL).
 * Means are unweighted over the selected points, H = sum m w (x-xbar)(y-ybar)^T, R = V U^T with the
 * det<0 reflection fix; writes the reference's return values R_out = R^T [9], t_out = -R^T t [3].
 * comp_R [9] / comp_t [3] (may be NULL) are updated in place as comp_R <- R_out comp_R,
 * comp_t <- R_out comp_t + t_out (losses.py:463-465). */
size_t rslo_kabsch_workspace_bytes(void);
int rslo_kabsch(const float* src, const float* tgt, const int32_t* tgt_idx, const float* weight,
                const float* normal, const float* mask, const float* dist, const float* dist_threshold, int n,
                float* R_out, float* t_out, float* comp_R, float* comp_t, void* workspace, size_t workspace_bytes,
                rslo_stream_t stream);

/* ROI threshold (losses.py:326-334): out[0] = max(k-th smallest of values[0..n), floor_value), k 1-based;
 * replaces torch.kthvalue + torch.max and keeps the threshold on the device. */
int rslo_kth_threshold(const float* values, int n, int k, float floor_value, float* out, rslo_stream_t stream);

/* ---- a11: covariance-weighted residual of the consistency loss (losses.py:348-363, 401-435) --------
 * pred [n,3], target [m,3], idx [n] (association into target), cov_pred [n,7], cov_target [m,7] raw
 * covariance parameters (3 eigenvalue increments + quaternion x,y,z,w), R [9] detached predicted rotation,
 * ROI = dist[i] < *dist_threshold.  loss = mean_roi(d^T S^-1 d) + reg * mean_roi(0.5 log det S),
 * S = C(cov_pred[i]) + R C(cov_target[idx[i]]) R^T.  sums: 4 doubles of scratch that the backward reads
 * (sum, logdet sum, ROI count).  Backward writes grad_pred [n,3] (may be NULL), grad_cov_pred [n,7] and
 * accumulates grad_target [m,3], grad_cov_target [m,7] (zeroed inside). */
int rslo_cov_residual_forward(const float* pred, const float* target, const int32_t* idx, const float* cov_pred,
                              const float* cov_target, const float* R, const float* dist,
                              const float* dist_threshold, int n, float reg_weight, double* sums, float* loss,
                              rslo_stream_t stream);
int rslo_cov_residual_backward(const float* pred, const float* target, const int32_t* idx, const float* cov_pred,
                               const float* cov_target, const float* R, const float* dist,
                               const float* dist_threshold, int n, int m, float reg_weight, const double* sums,
                               const float* grad_loss, float* grad_pred, float* grad_target, float* grad_cov_pred,
                               float* grad_cov_target, rslo_stream_t stream);

/* ---- a6 (covariance decoder): BatchNorm1d + LeakyReLU over stacked frames (csrc/bn1d_seg.cu) ---------------
 * Replaces nn.BatchNorm1d(C) + nn.LeakyReLU on `.features` after the decoder's convolutions
 * (rslo/models/middle.py:181-213).  x, z [N, C] rows of G stacked frames, frame g = seg_rows_host[g] consecutive rows
 * (G <= 16, C a multiple of 4).  training != 0: every frame is normalised with its own batch statistics (biased
 * variance) and running_mean / running_var (may be NULL) move once per non-empty frame, in frame order (momentum,
 * unbiased variance), num_batches_tracked += frames; stats: double [G][C][2] scratch ZEROED by the caller.
 * training == 0: running statistics.  slope >= 0: LeakyReLU(slope) follows; slope < 0: no activation.
 * mean_rstd: float [G][C][2] saved for the backward.
 * Backward: dz, x -> dx, dgamma [C], dbeta [C] (summed over the frames); sums: double [G][C][2], ZEROED by the caller. */
int rslo_bn1d_seg_forward(const float* x, int C, const int* seg_rows_host, int G, const float* gamma, const float* beta,
                          float* running_mean, float* running_var, long long* num_batches_tracked, float eps,
                          float momentum, int training, float slope, double* stats, float* z, float* mean_rstd,
                          rslo_stream_t stream);
int rslo_bn1d_seg_backward(const float* dz, const float* x, int C, const int* seg_rows_host, int G, const float* mean_rstd,
                           const float* gamma, const float* beta, float slope, int batch_stats, double* sums, float* dx,
                           float* dgamma, float* dbeta, rslo_stream_t stream);

/* ---- a13 glue: predicted pose applied to the target frame (csrc/pair_transform.cu) -----------------------------
 * y [n,3] = x @ R(q)^T + t with kornia 0.4.0's quaternion_to_rotation_matrix (q given as (w,x,y,z), normalised with
 * eps 1e-12), replacing the torch chain at rslo/models/voxel_odom_net.py:671-690; x rows are ldx floats apart (the
 * xyz columns of the voxel features).  identity != 0: R = I, t = 0 (global step <= 1500, :677-679).  R_out [9]
 * receives R.  Backward: grad_y, x -> dq (w,x,y,z) [4], dt [3]; workspace: rslo_pair_transform_workspace_bytes(),
 * ZEROED once by the caller (left zeroed). */
size_t rslo_pair_transform_workspace_bytes(void);
int rslo_pair_transform_forward(const float* x, int ldx, int n, const float* q_wxyz, const float* t, int identity, float* y,
                                float* R_out, rslo_stream_t stream);
int rslo_pair_transform_backward(const float* grad_y, const float* x, int ldx, int n, const float* q_wxyz, float* dq_wxyz,
                                 float* dt, void* workspace, size_t workspace_bytes, rslo_stream_t stream);

/* ---- f-N4: KITTI odometry sequence evaluation (csrc/kitti_eval.cu), float64 -----------------------------------
 * rslo_odom_to_abs_pose: relative poses [n,7] (t, q = w,x,y,z) -> absolute poses [n,7] exactly as
 * geometric.odom_to_abs_pose chains them (rslo/utils/geometric.py:376-406: abs[0] = identity, running pose starts at
 * odoms[0], quaternion renormalised with eps 1e-6 every step); two sequences at once (either may be NULL); dist_b
 * (may be NULL) [n] = cumulative trajectory length of sequence b (kittiOdomEval.trajectoryDistances,
 * rslo/utils/kitti_evaluation.py:42-61).
 * rslo_kitti_sequence_errors: kittiOdomEval.calcSequenceErrors (:95-145): for every start frame 0, step, 2 step, ...
 * of the ground truth and the 8 segment lengths 100..800 m, row r = (start index) * 8 + (length index):
 * err[r] = {first_frame, r_err/len, t_err/len, len, speed}, valid[r] = 0 where the reference `continue`s. */
int rslo_odom_to_abs_pose(const double* odom_a, const double* odom_b, int n, double* abs_a, double* abs_b, double* dist_b,
                          rslo_stream_t stream);
int rslo_kitti_sequence_errors(const double* abs_pred, int n_pred, const double* abs_gt, int n_gt, const double* dist_gt,
                               int step, double* err, int32_t* valid, rslo_stream_t stream);

/* ---- f-N3: point-cloud normal estimation (csrc/normals.cu) -------------------------------------------------------
 * Replaces estimate_normal() of script/create_hdf5.py:130-147 (open3d estimate_normals with
 * KDTreeSearchParamHybrid(radius, max_nn) + orient_normals_towards_camera_location): xyz rows are ld floats apart;
 * neighbours = the <= max_nn (<= 32) nearest points within radius, the point itself included; >= 3 neighbours: unit
 * eigenvector of the smallest eigenvalue of their covariance, else (0,0,1); every normal flipped to face camera_host[3].
 * normals [n,3].  workspace: rslo_estimate_normals_workspace_bytes(n). */
size_t rslo_estimate_normals_workspace_bytes(int n);
int rslo_estimate_normals(const float* xyz, int ld, int n, float radius, int max_nn, const float* camera_host,
                          float* normals, void* workspace, size_t workspace_bytes, rslo_stream_t stream);

/* ---- f-N2: the optimizer step in two launches (csrc/optim.cu) --------------------------------------------
 * Replaces torch.nn.utils.clip_grad_norm_(net.parameters(), 10.0) (train_hdf5.py:671) followed by
 * OptimWrapper.step() (rslo/torchplus/train/fastai_optim.py:181-194: p *= 1 - wd*lr on every trainable parameter,
 * then torch.optim.Adam(betas=(mom, 0.99)).step(), rslo/builder/optimizer_builder.py:101-118) over the flat gradient
 * buffer the all-reduce already uses.
 * rslo_grad_sumsq: *sumsq_out = sum of squares of grad[0..n) in double, bit-reproducible (fixed combine order).
 *   workspace: rslo_grad_norm_workspace_bytes() bytes, ZEROED once by the caller (the kernel leaves it zeroed).
 * rslo_adam_step: one CTA per chunk; chunk = up to a few thousand contiguous elements of ONE parameter:
 *   p = first element of the chunk in the parameter's own storage, off = its offset in grad / exp_avg / exp_avg_sq,
 *   flags bit 0 = the parameter received a gradient this step (otherwise only the weight decay is applied, as
 *   torch's Adam skips parameters whose .grad is None).
 *   g' = grad_scale * grad (grad_scale = 1 / world_size folds the all-reduce average in); when sumsq != NULL and
 *   max_norm > 0 the clip coefficient min(1, max_norm / (grad_scale * sqrt(*sumsq) + 1e-6)) multiplies g';
 *   true_wd != 0: p *= 1 - weight_decay * lr, else g' += weight_decay * p (Adam's L2);
 *   m = m + (1-beta1)(g'-m); v = beta2 v + (1-beta2) g'^2; p -= lr/(1-beta1^step) * m / (sqrt(v)/sqrt(1-beta2^step) + eps).
 *   write_clipped_grad != 0 stores g' back into grad (what clip_grad_norm_ leaves behind). */
typedef struct {
    float* p;
    unsigned int off;
    unsigned int n;
    unsigned int flags;
    unsigned int reserved;
} rslo_adam_chunk_t;
size_t rslo_grad_norm_workspace_bytes(void);
int rslo_grad_sumsq(const float* grad, size_t n, double* sumsq_out, void* workspace, size_t workspace_bytes,
                    rslo_stream_t stream);
int rslo_adam_step(const rslo_adam_chunk_t* chunks_dev, int n_chunks, float* grad, float* exp_avg, float* exp_avg_sq,
                   const double* sumsq, float grad_scale, float max_norm, float lr, float beta1, float beta2, float eps,
                   float weight_decay, int true_wd, int step, int write_clipped_grad, rslo_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RSLO_B200_H */
