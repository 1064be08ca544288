"""Torch-tensor wrappers over the C ABI (include/rslo_b200.h).

Every function here launches hand-written sm_100a kernels on the current CUDA stream through
``rslo_b200._lib``; nothing falls back to PyTorch or the CPU.  ``LAUNCHES`` counts kernel-launching
ABI calls (bench.py reports it).
"""
import ctypes as C

import numpy as np
import torch

from ._lib import check, lib, ptr, raw_stream, stream, workspace

LAUNCHES = {"calls": 0}
PROFILE = None      # set to a list to record (name, start_event, end_event, algorithmic_bytes, flops) per call


def _count(n=1):
    LAUNCHES["calls"] += n


def kernel_launch_count():
    """Kernels launched by librslo_b200 in this process (counted at the launch sites in csrc/)."""
    return int(lib.rslo_kernel_launch_count())


def _valid_entries(nbr):
    """Rulebook size R of a table (cached on the tensor; profiling pass only)."""
    r = getattr(nbr, "_rslo_R", None)
    if r is None:
        r = int((nbr >= 0).sum().item())
        try:
            nbr._rslo_R = r
        except AttributeError:
            pass
    return r


def _profiled(name, cost):
    """When PROFILE is a list, bracket the call with CUDA events on the current stream and record
    its algorithmic bytes / flops (SURVEY.md §8d formulas)."""
    def deco(fn):
        def wrapped(*a, **k):
            if PROFILE is None:
                return fn(*a, **k)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            out = fn(*a, **k)
            e.record()
            nbytes, flops = cost(out, *a, **k)
            PROFILE.append((name, s, e, nbytes, flops))
            return out
        wrapped.__name__, wrapped.__doc__ = fn.__name__, fn.__doc__
        return wrapped
    return deco


def _cost_spconv(out, feat, nbr, n_out, weight, *a, **k):
    Kk = nbr.shape[1]
    cin, cout = weight.shape[-2], weight.shape[-1]
    R = _valid_entries(nbr)
    return 4 * (feat.shape[0] * cin + n_out * cout + Kk * cin * cout) + 8 * R, 2 * R * cin * cout


def _cost_spconv_bwd_data(out, grad_out, nbr_t, n_in, weight, mirror):
    Kk = nbr_t.shape[1]
    cin, cout = weight.shape[-2], weight.shape[-1]
    R = _valid_entries(nbr_t)
    return 4 * (grad_out.shape[0] * cout + n_in * cin + Kk * cin * cout) + 8 * R, 2 * R * cin * cout


def _cost_spconv_bwd_weight(out, feat, grad_out, nbr, n_out, weight_shape, need_bias=True):
    Kk = nbr.shape[1]
    cin, cout = weight_shape[-2], weight_shape[-1]
    R = _valid_entries(nbr)
    return 4 * (feat.shape[0] * cin + n_out * cout + Kk * cin * cout) + 8 * R, 2 * R * cin * cout


def _i32(t):
    assert t.dtype == torch.int32 and t.is_cuda and t.is_contiguous(), "expect contiguous cuda int32"
    return t


def _f32(t):
    assert t.dtype == torch.float32 and t.is_cuda and t.is_contiguous(), "expect contiguous cuda float32"
    return t


def _nwords(D, H, W):
    return (D * H * W + 31) // 32


# ------------------------------------------------------------------------------------------------
# a10: nearest neighbour
# ------------------------------------------------------------------------------------------------
@_profiled("nn_exact", lambda out, q, t, brute=False: (12 * (q.shape[0] + t.shape[0]) + 8 * q.shape[0], 0))
def nn_exact(query, target, brute=False):
    """query [n,3], target [m,3] f32 cuda -> (dist [n] f32, idx [n] i32).  Bit-identical to the
    reference ChamferDistanceKernel (thirdparty/chamfer_distance/chamfer_distance.cu:6-137)."""
    q = _f32(query.contiguous())
    t = _f32(target.contiguous())
    n, m = q.shape[0], t.shape[0]
    dist = torch.empty(n, dtype=torch.float32, device=q.device)
    idx = torch.empty(n, dtype=torch.int32, device=q.device)
    if n == 0:
        return dist, idx
    if brute:
        check(lib.rslo_nn_brute(ptr(q), n, ptr(t), m, ptr(dist), ptr(idx), stream()), "rslo_nn_brute")
    else:
        nb = lib.rslo_nn_workspace_bytes(n, m)
        ws = workspace(nb, "nn")
        check(lib.rslo_nn_exact(ptr(q), n, ptr(t), m, ptr(dist), ptr(idx), ptr(ws), ws.numel(), stream()),
              "rslo_nn_exact")
    _count()
    return dist, idx


# ------------------------------------------------------------------------------------------------
# site tables / rulebooks
# ------------------------------------------------------------------------------------------------
class SiteTable:
    """Bitmap + popcount-prefix index of a level's active sites (see csrc/rulebook.cu)."""

    def __init__(self, shape, cells, perm):
        self.shape = tuple(int(s) for s in shape)    # (D, H, W)
        self.cells = cells                           # int32 [2*nwords]
        self.perm = perm                             # int32 [n] or None (rows already sorted)


def site_table_build(coors, n, shape, n_dev=None, need_perm=True):
    D, H, W = (int(s) for s in shape)
    coors = _i32(coors)
    cells = torch.empty(2 * _nwords(D, H, W), dtype=torch.int32, device=coors.device)
    perm = torch.empty(max(n, 1), dtype=torch.int32, device=coors.device) if need_perm else None
    nb = lib.rslo_site_table_workspace_bytes(D, H, W)
    ws = workspace(nb, "rulebook")
    check(lib.rslo_site_table_build(ptr(coors), coors.shape[1], n, ptr(n_dev), D, H, W, ptr(cells), ptr(perm),
                                    ptr(ws), ws.numel(), stream()), "rslo_site_table_build")
    _count()
    return SiteTable((D, H, W), cells, perm)


@_profiled("subm_table", lambda out, coors, n, table, ksize=(3, 3, 3), n_dev=None:
           (16 * n + 4 * n * int(np.prod(ksize)) + 8 * _nwords(*table.shape), 0))
def subm_table(coors, n, table, ksize=(3, 3, 3), n_dev=None):
    coors = _i32(coors)
    D, H, W = table.shape
    K = int(np.prod(ksize))
    nbr = torch.empty((max(n, 1), K), dtype=torch.int32, device=coors.device)
    check(lib.rslo_subm_table(ptr(coors), coors.shape[1], n, ptr(n_dev), D, H, W, ptr(table.cells),
                              ptr(table.perm), ksize[0], ksize[1], ksize[2], ptr(nbr), stream()),
          "rslo_subm_table")
    _count()
    return nbr


def out_shape_of(shape, ksize, stride, pad):
    return tuple((int(s) + 2 * p - k) // st + 1 for s, k, st, p in zip(shape, ksize, stride, pad))


@_profiled("strided_table", lambda out, coors, n, shape, ksize, *a, **k:
           (16 * n + 16 * out[1].shape[0] + 4 * (n + out[1].shape[0]) * int(np.prod(ksize)), 0))
def strided_table(coors, n, shape, ksize, stride, pad, n_dev=None, out_cap=None):
    """-> (out_table, out_coors [cap,4], n_out_dev [2] i32 {clamped, raw}, nbr [cap,K], nbr_inv [n,K])."""
    coors = _i32(coors)
    D, H, W = (int(s) for s in shape)
    oD, oH, oW = out_shape_of(shape, ksize, stride, pad)
    K = int(np.prod(ksize))
    if out_cap is None:
        reach = int(np.prod([-(-k // s) for k, s in zip(ksize, stride)]))
        out_cap = max(1, min(n * reach, oD * oH * oW))
    dev = coors.device
    cells = torch.empty(2 * _nwords(oD, oH, oW), dtype=torch.int32, device=dev)
    out_coors = torch.empty((out_cap, 4), dtype=torch.int32, device=dev)
    n_out_dev = torch.empty(2, dtype=torch.int32, device=dev)
    nbr = torch.empty((out_cap, K), dtype=torch.int32, device=dev)
    nbr_inv = torch.empty((max(n, 1), K), dtype=torch.int32, device=dev)
    nb = lib.rslo_strided_workspace_bytes(oD, oH, oW)
    ws = workspace(nb, "rulebook")
    check(lib.rslo_strided_table(ptr(coors), coors.shape[1], n, ptr(n_dev), D, H, W, ksize[0], ksize[1],
                                 ksize[2], stride[0], stride[1], stride[2], pad[0], pad[1], pad[2],
                                 ptr(cells), ptr(out_coors), out_cap, ptr(n_out_dev), ptr(nbr), ptr(nbr_inv),
                                 ptr(ws), ws.numel(), stream()), "rslo_strided_table")
    _count()
    return SiteTable((oD, oH, oW), cells, None), out_coors, n_out_dev, nbr, nbr_inv


def table_concat_many(jobs):
    """jobs: list of (tables, rows, adds) as for table_concat -> list of concatenated tables, ONE launch for all of them
    (the 12 rulebooks x T frames of a step; csrc/rulebook.cu:k_table_concat_multi)."""
    from ._lib import ConcatSeg
    outs, segs = [], []
    for tables, rows, adds in jobs:
        Kk = tables[0].shape[1]
        total = int(sum(rows))
        out = torch.empty((max(total, 1), Kk), dtype=torch.int32, device=tables[0].device)
        base = out.data_ptr()
        r0 = 0
        for t, r, a in zip(tables, rows, adds):
            if r > 0:
                segs.append((_i32(t).data_ptr(), base + 4 * r0 * Kk, int(r) * Kk, int(a)))
            r0 += r
        outs.append(out)
    if segs:
        arr = (ConcatSeg * len(segs))()
        for i, (sp, dp, cnt, add) in enumerate(segs):
            e = arr[i]
            e.src, e.dst, e.count, e.add = sp, dp, cnt, add
        check(lib.rslo_table_concat_multi(arr, len(segs), stream()), "rslo_table_concat_multi")
        _count()
    return outs


def table_concat(tables, rows, adds):
    """Row-wise concatenation of per-frame tables [rows_f, K] with row indices shifted by adds[f]."""
    Kk = tables[0].shape[1]
    total = int(sum(rows))
    out = torch.empty((max(total, 1), Kk), dtype=torch.int32, device=tables[0].device)
    r0 = 0
    for t, r, a in zip(tables, rows, adds):
        if r > 0:
            check(lib.rslo_table_concat(ptr(_i32(t)), int(r) * Kk, int(a), ptr(out[r0:r0 + r]), stream()),
                  "rslo_table_concat")
            _count()
        r0 += r
    return out


# ------------------------------------------------------------------------------------------------
# voxeliser
# ------------------------------------------------------------------------------------------------
@_profiled("voxelize", lambda out, points, *a, **k:
           (points.numel() * 4 + (28 + 16) * out["coordinates"].shape[0] +
            (0 if out["voxels"] is None else out["voxels"].numel() * 4), 0))
def voxelize(points, voxel_size, pc_range, grid_size, max_points=10, max_voxels=40000, block_factor=1,
             block_size=8, height_threshold=-1.0, batch_idx=0, materialize=True, with_mean=True,
             with_table=False, coor_stride=4):
    """points [P,F] f32 cuda -> dict of capacity-sized device tensors + 'n_dev' (device count).

    Replaces `_VoxelGenerator.generate` (rslo/builder/voxel_builder.py:48-54) and, with
    ``with_mean``, `SimpleVoxel_XYZINormalC.forward` (rslo/models/voxel_encoder.py:272-280)."""
    pts = _f32(points.contiguous())
    P, F = pts.shape
    gx, gy, gz = (int(g) for g in grid_size)
    dev = pts.device
    table_d = gz + 1
    voxels = torch.empty((max_voxels, max_points, F), dtype=torch.float32, device=dev) if materialize else None
    coors = torch.empty((max_voxels, coor_stride), dtype=torch.int32, device=dev)
    num = torch.empty(max_voxels, dtype=torch.int32, device=dev)
    mean = torch.empty((max_voxels, 7), dtype=torch.float32, device=dev) if with_mean else None
    n_dev = torch.empty(1, dtype=torch.int32, device=dev)
    cells = perm = None
    if with_table:
        cells = torch.empty(2 * _nwords(table_d, gy, gx), dtype=torch.int32, device=dev)
        perm = torch.empty(max(P, 1), dtype=torch.int32, device=dev)
    vs = (C.c_float * 3)(*[float(v) for v in voxel_size])
    rg = (C.c_float * 6)(*[float(v) for v in pc_range])
    nb = lib.rslo_voxelize_workspace_bytes(P, gx, gy, table_d, max_voxels)
    ws = workspace(nb, "voxelize")
    check(lib.rslo_voxelize(ptr(pts), P, F, vs, rg, gx, gy, gz, max_points, max_voxels, block_factor,
                            block_size, float(height_threshold), batch_idx, ptr(voxels), ptr(coors),
                            coor_stride, ptr(num), ptr(mean), ptr(n_dev), ptr(cells), ptr(perm), table_d,
                            ptr(ws), ws.numel(), stream()), "rslo_voxelize")
    _count()
    out = {"voxels": voxels, "coordinates": coors, "num_points_per_voxel": num, "mean": mean, "n_dev": n_dev}
    if with_table:
        out["table"] = SiteTable((table_d, gy, gx), cells, perm)
    return out


def vfe_mean(voxels, num_points):
    v = _f32(voxels.contiguous())
    npts = _i32(num_points.contiguous())
    n, mp, F = v.shape
    mean = torch.empty((n, 7), dtype=torch.float32, device=v.device)
    check(lib.rslo_vfe_mean(ptr(v), ptr(npts), n, mp, F, ptr(mean), stream()), "rslo_vfe_mean")
    _count()
    return mean


# ------------------------------------------------------------------------------------------------
# sparse convolution
# ------------------------------------------------------------------------------------------------
@_profiled("spconv_forward", _cost_spconv)
def spconv_forward(feat, nbr, n_out, weight, bias=None, scale=None, shift=None, act=0, slope=0.01,
                   n_out_dev=None):
    feat = _f32(feat)
    nbr = _i32(nbr)
    K = nbr.shape[1]
    w = _f32(weight.contiguous())
    Cin, Cout = w.shape[-2], w.shape[-1]
    assert w.numel() == K * Cin * Cout and feat.shape[1] == Cin
    out = torch.empty((n_out, Cout), dtype=torch.float32, device=feat.device)
    if n_out == 0:
        return out
    check(lib.rslo_spconv_forward(ptr(feat), ptr(nbr), n_out, ptr(n_out_dev), K, Cin, Cout, ptr(w),
                                  ptr(bias), ptr(scale), ptr(shift), act, float(slope), ptr(out), stream()),
          "rslo_spconv_forward")
    _count()
    return out


def spconv_tc_supported(Cin, Cout, K):
    return bool(lib.rslo_spconv_tc_supported(int(Cin), int(Cout), int(K)))


def spconv_tc_prepare(weight, transpose=False, mirror=False):
    """W [K,Cin,Cout] -> pre-swizzled 3xTF32 {hi, lo} image for csrc/spconv_tc.cu."""
    w = _f32(weight.contiguous())
    Kk, Cin, Cout = w.shape[-3] if w.dim() == 3 else w.numel() // (w.shape[-2] * w.shape[-1]), w.shape[-2], w.shape[-1]
    img = torch.empty(lib.rslo_spconv_tc_image_bytes(Kk, Cin, Cout) // 4, dtype=torch.float32, device=w.device)
    check(lib.rslo_spconv_tc_prepare(ptr(w), Kk, Cin, Cout, 1 if transpose else 0, 1 if mirror else 0, ptr(img),
                                     stream()), "rslo_spconv_tc_prepare")
    _count()
    return img


def _cost_spconv_tc(out, feat, nbr, n_out, image, kdim, ndim, *a, **k):
    Kk = nbr.shape[1]
    R = _valid_entries(nbr)
    return 4 * (feat.shape[0] * kdim + n_out * ndim + Kk * kdim * ndim) + 8 * R, 2 * R * kdim * ndim


@_profiled("spconv_tc", _cost_spconv_tc)
def spconv_tc_forward(feat, nbr, n_out, image, kdim, ndim, bias=None, act=0, slope=0.01, n_out_dev=None):
    """out[o,:] = act(bias + sum_k feat[nbr[o,k],:] @ B_k) on tcgen05 tensor cores (3xTF32)."""
    feat = _f32(feat)
    nbr = _i32(nbr)
    assert feat.shape[1] == kdim
    out = torch.empty((n_out, ndim), dtype=torch.float32, device=feat.device)
    if n_out == 0:
        return out
    ws = workspace(lib.rslo_spconv_tc_workspace_bytes(n_out, ndim), "spconv_tc")
    check(lib.rslo_spconv_tc_forward(ptr(feat), ptr(nbr), n_out, ptr(n_out_dev), nbr.shape[1], kdim, ndim, ptr(image),
                                     ptr(bias), act, float(slope), ptr(out), ptr(ws), ws.numel(), stream()),
          "rslo_spconv_tc_forward")
    _count()
    return out


@_profiled("spconv_backward_data", _cost_spconv_bwd_data)
def spconv_backward_data(grad_out, nbr_t, n_in, weight, mirror):
    g = _f32(grad_out.contiguous())
    nbr_t = _i32(nbr_t)
    K = nbr_t.shape[1]
    w = _f32(weight.contiguous())
    Cin, Cout = w.shape[-2], w.shape[-1]
    wt = torch.empty(K * Cin * Cout, dtype=torch.float32, device=g.device)
    check(lib.rslo_spconv_transpose_weight(ptr(w), K, Cin, Cout, 1 if mirror else 0, ptr(wt), stream()),
          "rslo_spconv_transpose_weight")
    grad_in = torch.empty((n_in, Cin), dtype=torch.float32, device=g.device)
    if n_in > 0:
        check(lib.rslo_spconv_backward_data(ptr(g), ptr(nbr_t), n_in, None, K, Cin, Cout, ptr(wt),
                                            ptr(grad_in), stream()), "rslo_spconv_backward_data")
    _count(2)
    return grad_in


@_profiled("spconv_backward_weight", _cost_spconv_bwd_weight)
def spconv_backward_weight(feat, grad_out, nbr, n_out, weight_shape, need_bias=True):
    feat = _f32(feat)
    g = _f32(grad_out.contiguous())
    nbr = _i32(nbr)
    K = nbr.shape[1]
    Cin, Cout = weight_shape[-2], weight_shape[-1]
    gw = torch.empty(tuple(weight_shape), dtype=torch.float32, device=g.device)
    gb = torch.empty(Cout, dtype=torch.float32, device=g.device) if need_bias else None
    check(lib.rslo_spconv_backward_weight(ptr(feat), ptr(g), ptr(nbr), n_out, None, K, Cin, Cout,
                                          ptr(gw), ptr(gb), stream()), "rslo_spconv_backward_weight")
    _count(2)
    return gw, gb


def spconv_tc_wgrad_supported(Cin, Cout):
    return bool(lib.rslo_spconv_tc_wgrad_supported(int(Cin), int(Cout)))


@_profiled("spconv_tc_wgrad", _cost_spconv_bwd_weight)
def spconv_tc_backward_weight(feat, grad_out, nbr, n_out, weight_shape, need_bias=False):
    """dW on the tcgen05 tensor cores (split-TF32); Cin, Cout in {32, 64}."""
    feat = _f32(feat)
    g = _f32(grad_out.contiguous())
    nbr = _i32(nbr)
    Cin, Cout = weight_shape[-2], weight_shape[-1]
    gw = torch.empty(tuple(weight_shape), dtype=torch.float32, device=g.device)
    check(lib.rslo_spconv_tc_backward_weight(ptr(feat), ptr(g), ptr(nbr), n_out, None, nbr.shape[1], Cin, Cout,
                                             ptr(gw), stream()), "rslo_spconv_tc_backward_weight")
    _count()
    return gw


def act_backward(grad_out, out, act, slope, need_bias=True):
    """LeakyReLU backward from the saved output fused with the bias gradient -> (grad_act, grad_bias)."""
    g = _f32(grad_out.contiguous())
    n, Cc = g.shape
    ga = torch.empty_like(g) if act else g
    gb = torch.empty(Cc, dtype=torch.float32, device=g.device) if need_bias else None
    check(lib.rslo_act_backward(ptr(g), ptr(out), n, None, Cc, int(act), float(slope), ptr(ga) if act else None,
                                ptr(gb), stream()), "rslo_act_backward")
    _count()
    return ga, gb


@_profiled("dense_from_sites", lambda out, feat, table: (feat.numel() * 4 + out.numel() * 4, 0))
def dense_from_sites(feat, table):
    feat = _f32(feat)
    D, H, W = table.shape
    Cc = feat.shape[1]
    dense = torch.empty((Cc * D, H, W), dtype=torch.float32, device=feat.device)
    check(lib.rslo_dense_from_sites(ptr(feat), Cc, ptr(table.cells), ptr(table.perm), D, H, W, ptr(dense),
                                    stream()), "rslo_dense_from_sites")
    _count()
    return dense


def dense_backward(grad_dense, coors, n, shape, C_):
    g = _f32(grad_dense.contiguous())
    coors = _i32(coors)
    D, H, W = shape
    gf = torch.empty((n, C_), dtype=torch.float32, device=g.device)
    if n > 0:
        check(lib.rslo_dense_backward(ptr(g), C_, ptr(coors), coors.shape[1], n, None, D, H, W, ptr(gf),
                                      stream()), "rslo_dense_backward")
    _count()
    return gf


# ------------------------------------------------------------------------------------------------
# a12: weighted Kabsch
# ------------------------------------------------------------------------------------------------
@_profiled("kabsch", lambda out, src, *a, **k: (src.shape[0] * 4 * (3 + 3 + 1 + 1), 0))
def kabsch(src, tgt, weight=None, mask=None, dist=None, dist_threshold=None, comp_R=None, comp_t=None, tgt_idx=None,
           normal=None):
    """src [n,3], tgt rows -> (R [3,3], t [3]) with SVDHead's return convention (rslo/layers/svd.py:57-64).
    Sync-free: selection by `mask` and/or `dist < dist_threshold`, the association gather (`tgt_idx`) and
    the normal-cosine weight (`normal`) all happen inside the reduction."""
    src = _f32(src.contiguous())
    tgt = _f32(tgt.contiguous())
    n = src.shape[0]
    R = torch.empty((3, 3), dtype=torch.float32, device=src.device)
    t = torch.empty(3, dtype=torch.float32, device=src.device)
    ws = workspace(lib.rslo_kabsch_workspace_bytes(), "kabsch")
    check(lib.rslo_kabsch(ptr(src), ptr(tgt), ptr(tgt_idx), ptr(weight), ptr(normal), ptr(mask), ptr(dist),
                          ptr(dist_threshold), n, ptr(R), ptr(t), ptr(comp_R), ptr(comp_t), ptr(ws), ws.numel(),
                          stream()), "rslo_kabsch")
    _count()
    return R, t


def kth_threshold(values, k, floor_value):
    """max(k-th smallest of values, floor_value) -> [1] on the device (losses.py:326-334), no host sync."""
    v = _f32(values.contiguous().reshape(-1))
    out = torch.empty(1, dtype=torch.float32, device=v.device)
    check(lib.rslo_kth_threshold(ptr(v), v.numel(), int(k), float(floor_value), ptr(out), stream()),
          "rslo_kth_threshold")
    _count()
    return out


# ------------------------------------------------------------------------------------------------
# a11: covariance-weighted residual (fused forward / backward)
# ------------------------------------------------------------------------------------------------
class _CovResidualFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, cov_pred, cov_target, R, idx, dist, thr, reg):
        pred, target = _f32(pred.contiguous()), _f32(target.contiguous())
        cov_pred, cov_target = _f32(cov_pred.contiguous()), _f32(cov_target.contiguous())
        R = _f32(R.contiguous())
        n = pred.shape[0]
        sums = torch.empty(4, dtype=torch.float64, device=pred.device)
        loss = torch.empty(1, dtype=torch.float32, device=pred.device)
        check(lib.rslo_cov_residual_forward(ptr(pred), ptr(target), ptr(idx), ptr(cov_pred), ptr(cov_target), ptr(R),
                                            ptr(dist), ptr(thr), n, float(reg), ptr(sums), ptr(loss), stream()),
              "rslo_cov_residual_forward")
        _count()
        ctx.save_for_backward(pred, target, cov_pred, cov_target, R, idx, dist, thr, sums)
        ctx.reg = float(reg)
        return loss

    @staticmethod
    def backward(ctx, g):
        pred, target, cov_pred, cov_target, R, idx, dist, thr, sums = ctx.saved_tensors
        g = _f32(g.contiguous())
        n, m = pred.shape[0], target.shape[0]
        g_pred = torch.empty_like(pred) if ctx.needs_input_grad[0] else None
        g_t = torch.empty_like(target)
        g_cp = torch.empty_like(cov_pred)
        g_ct = torch.empty((m, 7), dtype=torch.float32, device=pred.device)
        check(lib.rslo_cov_residual_backward(ptr(pred), ptr(target), ptr(idx), ptr(cov_pred), ptr(cov_target), ptr(R),
                                             ptr(dist), ptr(thr), n, m, ctx.reg, ptr(sums), ptr(g), ptr(g_pred),
                                             ptr(g_t), ptr(g_cp), ptr(g_ct), stream()), "rslo_cov_residual_backward")
        _count()
        return g_pred, g_t, g_cp, g_ct, None, None, None, None, None


def cov_residual(pred, target, cov_pred, cov_target, R, idx, dist, thr, reg_weight):
    """mean_roi(d^T S^-1 d) + reg * mean_roi(0.5 log det S) -> [1]; differentiable w.r.t. pred, target,
    cov_pred, cov_target (rslo/core/losses.py:348-363, 401-435)."""
    return _CovResidualFn.apply(pred, target, cov_pred, cov_target, R, _i32(idx), _f32(dist), _f32(thr.reshape(1)),
                                reg_weight)


# ------------------------------------------------------------------------------------------------
# a8: dense 2-D convolutions of the head (csrc/conv2d_tc.cu): TMA-staged split-TF32 tcgen05 implicit GEMM
# ------------------------------------------------------------------------------------------------
def conv2d_tc_supported(cin, cout, ksize, stride):
    return bool(lib.rslo_conv2d_tc_supported(cin, cout, ksize, stride))


_CONV_WS = {}


def _conv_ws():
    """Per-(device, stream) split-K workspace of the dense convolutions; zeroed once (the kernels keep the
    tile counters zeroed between launches)."""
    key = raw_stream()
    dev = key[0]
    ws = _CONV_WS.get(key)
    if ws is None:
        ws = torch.zeros(lib.rslo_conv2d_tc_workspace_bytes(0, 0, 0, 0), dtype=torch.uint8, device=f"cuda:{dev}")
        _CONV_WS[key] = ws
    return ws


def conv2d_split(x_nhwc, out=None):
    """x [B,H,W,C] f32 contiguous -> split pair [2,B,H,W,C] (hi = RN_tf32(x), lo = RN_tf32(x - hi))."""
    x = _f32(x_nhwc)
    if out is None:
        out = torch.empty((2,) + tuple(x.shape), dtype=torch.float32, device=x.device)
    check(lib.rslo_conv2d_split(ptr(x), x.numel(), ptr(out), stream()), "rslo_conv2d_split")
    _count()
    return out


def conv2d_tc_prepare(weight_oihw, mode, out=None, cout_padded=None):
    """OIHW weight -> split image [2, k*k, N, Kd] (mode 0 forward: N=Cout, Kd=Cin; mode 1 data gradient);
    `cout_padded` zero-pads the output channels (narrow 7- / 1-channel heads -> 32)."""
    w = _f32(weight_oihw.detach().contiguous())
    cout, cin, ks, _ = w.shape
    cp = cout if cout_padded is None else int(cout_padded)
    if out is None:
        out = torch.empty(2 * ks * ks * cin * cp, dtype=torch.float32, device=w.device)
    check(lib.rslo_conv2d_tc_prepare(ptr(w), cout, cp, cin, ks, mode, ptr(out), stream()), "rslo_conv2d_tc_prepare")
    _count()
    return out


def _cost_conv2d(res_, x_split, image, cout, ksize, stride, *a, **k):
    _, B, H, W, cin = x_split.shape
    npx = res_.shape[0] * res_.shape[1] * res_.shape[2]
    return 4 * (B * H * W * cin + npx * cout + ksize * ksize * cin * cout), 2 * npx * ksize * ksize * cin * cout


@_profiled("conv2d_tc", _cost_conv2d)
def conv2d_tc_forward(x_split, image, cout, ksize, stride, bias=None, relu=False, out=None, stats=None, imgs_per_group=1):
    """y [B,Ho,Wo,Cout] = conv(x) (+bias)(ReLU); x_split [2,B,H,W,Cin].  `stats` (float64 [G,Cout,2], zeroed by
    the caller) receives the per-group channel sums / sums of squares of y (BatchNorm batch statistics)."""
    x = _f32(x_split)
    _, B, H, W, cin = x.shape
    pad = ksize // 2
    Ho, Wo = (H + 2 * pad - ksize) // stride + 1, (W + 2 * pad - ksize) // stride + 1
    if out is None:
        out = torch.empty((B, Ho, Wo, cout), dtype=torch.float32, device=x.device)
    ws = _conv_ws()
    check(lib.rslo_conv2d_tc_forward(ptr(x), B, H, W, cin, ptr(image), cout, ksize, stride, ptr(bias), 1 if relu else 0,
                                     ptr(out), ptr(stats), imgs_per_group, ptr(ws), ws.numel(), stream()),
          "rslo_conv2d_tc_forward")
    _count()
    return out


def _cost_conv2d_bwd(res_, g_split, image_t, in_shape, ksize, stride, *a, **k):
    _, B, Ho, Wo, cout = g_split.shape
    _, H, W, cin = in_shape
    return 4 * (B * H * W * cin + B * Ho * Wo * cout + ksize * ksize * cin * cout), 2 * B * Ho * Wo * ksize * ksize * cin * cout


@_profiled("conv2d_tc", _cost_conv2d_bwd)          # same CUDA kernel (k_conv2d_tc) as the forward, other tap table
def conv2d_tc_backward_data(g_split, image_t, in_shape, ksize, stride, out=None, accumulate=False):
    """dx [B,H,W,Cin] (= or +=) from g_split [2,B,Ho,Wo,Cout] and the mode-1 weight image."""
    g = _f32(g_split)
    B, H, W, cin = in_shape
    cout = g.shape[-1]
    if out is None:
        assert not accumulate
        out = torch.empty((B, H, W, cin), dtype=torch.float32, device=g.device)
    ws = _conv_ws()
    check(lib.rslo_conv2d_tc_backward_data(ptr(g), B, H, W, cin, ptr(image_t), cout, ksize, stride, ptr(out),
                                           1 if accumulate else 0, ptr(ws), ws.numel(), stream()),
          "rslo_conv2d_tc_backward_data")
    _count()
    return out


def _cost_conv2d_wgrad(res_, x_split, g_split, ksize, stride, *a, **k):
    _, B, H, W, cin = x_split.shape
    _, _, Ho, Wo, cout = g_split.shape
    return 4 * (B * H * W * cin + B * Ho * Wo * cout + ksize * ksize * cin * cout), 2 * B * Ho * Wo * ksize * ksize * cin * cout


@_profiled("conv2d_tc_wgrad", _cost_conv2d_wgrad)
def conv2d_tc_backward_weight(x_split, g_split, ksize, stride, out=None, accumulate=False, cout_real=None, scratch=None):
    """grad_weight OIHW [Cout,Cin,k,k] from x_split [2,B,H,W,Cin] and g_split [2,B,Ho,Wo,Cout].
    `scratch` given (zeroed float32 [k*k*Cin*Cout]): the products are added into it as [k*k][Cin][Cout] and left
    there (no OIHW output; see conv2d_multi_wgrad_finish)."""
    x, g = _f32(x_split), _f32(g_split)
    _, B, H, W, cin = x.shape
    cout = g.shape[-1]
    cr = cout if cout_real is None else int(cout_real)
    deferred = scratch is not None
    if deferred:
        out = None
    else:
        if out is None:
            assert not accumulate
            out = torch.empty((cr, cin, ksize, ksize), dtype=torch.float32, device=x.device)
        scratch = workspace(lib.rslo_conv2d_tc_wgrad_scratch_bytes(cin, cout, ksize), "conv2d_wgrad")
    check(lib.rslo_conv2d_tc_backward_weight(ptr(x), ptr(g), B, H, W, cin, cout, ksize, stride, cr, ptr(scratch),
                                             1 if accumulate else 0, ptr(out), stream()),
          "rslo_conv2d_tc_backward_weight")
    _count()
    return out


# ------------------------------------------------------------------------------------------------
# a8: the head between its convolutions (csrc/head_ops.cu) — thin pointer-passing wrappers; the buffers are
# owned by rslo_b200/layers/head_tc.py
# ------------------------------------------------------------------------------------------------
def conv2d_multi_prepare(table_dev, n):
    check(lib.rslo_conv2d_multi_prepare(ptr(table_dev), n, stream()), "rslo_conv2d_multi_prepare")
    _count()


def conv2d_multi_wgrad_finish(table_dev, n):
    check(lib.rslo_conv2d_multi_wgrad_finish(ptr(table_dev), n, stream()), "rslo_conv2d_multi_wgrad_finish")
    _count()


def head_pack_input(x1, x2, split_out, mask_out):
    """x1, x2 [B,C,H,W] (NCHW) -> split pair [2,B,H,W,2C] of cat(x1,x2) and mask [B,H,W] = (sum_c x1 != 0)."""
    x1, x2 = _f32(x1), _f32(x2)
    B, Cc, H, W = x1.shape
    check(lib.rslo_head_pack_input(ptr(x1), ptr(x2), B, Cc, H * W, ptr(split_out), ptr(mask_out), stream()),
          "rslo_head_pack_input")
    _count()


def head_unpack_grad(dx, g1, g2):
    """dx [B,H,W,2C] -> g1, g2 [B,C,H,W]."""
    B, Cc, H, W = g1.shape
    check(lib.rslo_head_unpack_grad(ptr(dx), B, Cc, H * W, ptr(g1), ptr(g2), stream()), "rslo_head_unpack_grad")
    _count()


def bn_act_forward(y, ipg, stats, gamma, beta, running_mean, running_var, nbt, eps, momentum, update_repeat,
                   residual, relu, z, z_split, mean_rstd):
    B, H, W, Cc = y.shape
    check(lib.rslo_bn_act_forward(ptr(y), B, H * W, Cc, ipg, ptr(stats), ptr(gamma), ptr(beta), ptr(running_mean),
                                  ptr(running_var), ptr(nbt), float(eps), float(momentum), int(update_repeat),
                                  ptr(residual), 1 if relu else 0, ptr(z), ptr(z_split), ptr(mean_rstd), stream()),
          "rslo_bn_act_forward")
    _count()


def bn_act_backward(dz, z, y, ipg, mean_rstd, gamma, relu, batch_stats, sums, g_split, dres, dres_accumulate, dgamma,
                    dbeta, dbias):
    B, H, W, Cc = y.shape
    check(lib.rslo_bn_act_backward(ptr(dz), ptr(z), ptr(y), B, H * W, Cc, ipg, ptr(mean_rstd), ptr(gamma),
                                   1 if relu else 0, 1 if batch_stats else 0, ptr(sums), ptr(g_split), ptr(dres),
                                   1 if dres_accumulate else 0, ptr(dgamma), ptr(dbeta), ptr(dbias), stream()),
          "rslo_bn_act_backward")
    _count(2)


def upcat_split(z, up, ld, choff, dst_split):
    B, H, W, Cc = z.shape
    check(lib.rslo_upcat_split(ptr(z), B, H, W, Cc, up, ld, choff, ptr(dst_split), stream()), "rslo_upcat_split")
    _count()


def upcat_backward(dcat, shape, up, ld, choff, dz, accumulate):
    B, H, W, Cc = shape
    check(lib.rslo_upcat_backward(ptr(dcat), B, H, W, Cc, up, ld, choff, ptr(dz), 1 if accumulate else 0, stream()),
          "rslo_upcat_backward")
    _count()


def bias_grad(g, n_rows, ld, c, out):
    check(lib.rslo_bias_grad(ptr(g), n_rows, ld, c, ptr(out), stream()), "rslo_bias_grad")
    _count()


# ------------------------------------------------------------------------------------------------
# f-N2: optimizer step (csrc/optim.cu)
# ------------------------------------------------------------------------------------------------
def grad_norm_workspace_bytes():
    return int(lib.rslo_grad_norm_workspace_bytes())


def _cost_sumsq(res_, flat, *a, **k):
    return 4 * flat.numel(), 2 * flat.numel()


@_profiled("grad_sumsq", _cost_sumsq)
def grad_sumsq(flat, out, ws):
    """out[0] (float64) = sum of squares of the flat fp32 gradient buffer (bit-reproducible); ws: zeroed once."""
    assert flat.dtype == torch.float32 and flat.is_cuda and flat.is_contiguous() and out.dtype == torch.float64
    check(lib.rslo_grad_sumsq(ptr(flat), flat.numel(), ptr(out), ptr(ws), ws.numel(), stream()), "rslo_grad_sumsq")
    _count()
    return out


def _cost_adam(res_, table, n_chunks, flat, *a, **k):
    return 28 * flat.numel(), 12 * flat.numel()


@_profiled("adam_step", _cost_adam)
def adam_step(table, n_chunks, flat, exp_avg, exp_avg_sq, sumsq, grad_scale, max_norm, lr, beta1, beta2, eps,
              weight_decay, true_wd, step, write_clipped_grad=False):
    """clip coefficient + 1/world scaling + weight decay + Adam over every chunk of `table` (rslo_adam_chunk_t[])."""
    check(lib.rslo_adam_step(ptr(table), n_chunks, ptr(flat), ptr(exp_avg), ptr(exp_avg_sq), ptr(sumsq),
                             float(grad_scale), float(max_norm), float(lr), float(beta1), float(beta2), float(eps),
                             float(weight_decay), 1 if true_wd else 0, int(step), 1 if write_clipped_grad else 0,
                             stream()), "rslo_adam_step")
    _count()


# ------------------------------------------------------------------------------------------------
# covariance decoder: BatchNorm1d + LeakyReLU over stacked frames (csrc/bn1d_seg.cu)
# ------------------------------------------------------------------------------------------------
def _seg_array(seg):
    arr = (C.c_int * len(seg))(*[int(v) for v in seg])
    return arr


def _cost_bn1d(res_, x, *a, **k):
    return 12 * x.numel(), 0


@_profiled("bn1d_seg", _cost_bn1d)
def bn1d_seg_forward(x, seg, gamma, beta, running_mean, running_var, nbt, eps, momentum, training, slope):
    """x [N,C] rows of len(seg) stacked frames -> (z, mean_rstd [G,C,2]); see include/rslo_b200.h"""
    x = _f32(x)
    n, c = x.shape
    assert sum(int(v) for v in seg) == n
    G = len(seg)
    z = torch.empty_like(x)
    mean_rstd = torch.empty((G, c, 2), dtype=torch.float32, device=x.device)
    stats = torch.zeros((G, c, 2), dtype=torch.float64, device=x.device) if training else None
    check(lib.rslo_bn1d_seg_forward(ptr(x), c, _seg_array(seg), G, ptr(gamma), ptr(beta), ptr(running_mean),
                                    ptr(running_var), ptr(nbt), float(eps), float(momentum), 1 if training else 0,
                                    float(slope), ptr(stats), ptr(z), ptr(mean_rstd), stream()), "rslo_bn1d_seg_forward")
    _count(2 if training else 1)
    return z, mean_rstd


@_profiled("bn1d_seg", _cost_bn1d)
def bn1d_seg_backward(dz, x, seg, mean_rstd, gamma, beta, slope, batch_stats):
    """-> (dx [N,C], dgamma [C], dbeta [C])"""
    dz, x = _f32(dz), _f32(x)
    n, c = x.shape
    G = len(seg)
    dx = torch.empty_like(x)
    dgb = torch.empty((2, c), dtype=torch.float32, device=x.device)
    sums = torch.zeros((G, c, 2), dtype=torch.float64, device=x.device)
    check(lib.rslo_bn1d_seg_backward(ptr(dz), ptr(x), c, _seg_array(seg), G, ptr(mean_rstd), ptr(gamma), ptr(beta),
                                     float(slope), 1 if batch_stats else 0, ptr(sums), ptr(dx), ptr(dgb[0]), ptr(dgb[1]),
                                     stream()), "rslo_bn1d_seg_backward")
    _count(2)
    return dx, dgb[0], dgb[1]


# ------------------------------------------------------------------------------------------------
# a13 glue: predicted pose applied to the target frame's points (csrc/pair_transform.cu)
# ------------------------------------------------------------------------------------------------
_PX_WS = {}


class _PairTransformFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, q_wxyz, t, identity):
        # x: [n, >=3] rows (xyz first), may be a row-strided view of the voxel features
        assert x.dtype == torch.float32 and x.is_cuda and x.stride(1) == 1
        n, ldx = x.shape[0], x.stride(0)
        q = q_wxyz.detach().reshape(4).contiguous()
        tt = t.detach().reshape(3).contiguous()
        y = torch.empty((n, 3), dtype=torch.float32, device=x.device)
        R = torch.empty((3, 3), dtype=torch.float32, device=x.device)
        check(lib.rslo_pair_transform_forward(ptr(x), ldx, n, ptr(q), ptr(tt), 1 if identity else 0, ptr(y), ptr(R), stream()),
              "rslo_pair_transform_forward")
        _count()
        ctx.identity, ctx.qshape, ctx.tshape = identity, q_wxyz.shape, t.shape
        ctx.save_for_backward(x, q)
        ctx.mark_non_differentiable(R)
        return y, R

    @staticmethod
    def backward(ctx, gy, _gR):
        if ctx.identity:
            return None, None, None, None
        x, q = ctx.saved_tensors
        key = raw_stream()
        ws = _PX_WS.get(key)
        if ws is None:
            ws = _PX_WS[key] = torch.zeros(lib.rslo_pair_transform_workspace_bytes(), dtype=torch.uint8, device=x.device)
        out = torch.empty(7, dtype=torch.float32, device=x.device)
        gy = _f32(gy.contiguous())
        check(lib.rslo_pair_transform_backward(ptr(gy), ptr(x), x.stride(0), x.shape[0], ptr(q), ptr(out), ptr(out[4:]), ptr(ws),
                                               ws.numel(), stream()), "rslo_pair_transform_backward")
        _count()
        return None, out[:4].reshape(ctx.qshape), out[4:].reshape(ctx.tshape), None


def pair_transform(x, q_wxyz, t, identity=False):
    """(y [n,3] = x[:, :3] @ R(q)^T + t, R [3,3]); differentiable w.r.t. q (w,x,y,z) and t; x is treated as data
    (the reference detaches nothing here, but the points are encoder INPUTS without gradient)."""
    return _PairTransformFn.apply(x, q_wxyz, t, bool(identity))
