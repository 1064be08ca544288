"""Sparse 3-D convolution layers with the spconv 1.x module surface the reference uses
(`rslo/models/middle.py:80-97,119-213`): SparseConvTensor, SubMConv3d, SparseConv3d,
SparseInverseConv3d, SparseSequential.  Parameters keep the reference state_dict layout
(`weight [kD,kH,kW,Cin,Cout]`, `bias [Cout]`) so `ours.tckpt` loads unchanged.

Compute is the output-stationary gather-convolution kernel (csrc/spconv.cu) over index tables built
by csrc/rulebook.cu; there is no PyTorch or CPU fallback.
"""
import math
import os

import torch
from torch import nn
from torch.nn import functional as F

from .. import kernels as K


class IndexEntry:
    """Tables of one `indice_key`: forward table, its transpose, site sets on both sides."""

    def __init__(self, kind, nbr, nbr_t, n_in, n_out, out_indices, out_shape, out_table, mirror):
        self.kind = kind                # "subm" | "strided"
        self.nbr = nbr                  # [n_out, K] input row per offset or -1
        self.nbr_t = nbr_t              # table of the transposed operator
        self.n_in, self.n_out = n_in, n_out
        self.out_indices = out_indices  # [n_out,4] (b,z,y,x)
        self.out_shape = out_shape
        self.out_table = out_table      # kernels.SiteTable of the output level
        self.mirror = mirror            # transposed filter bank uses mirrored offsets (subm)
        self.seg_in = self.seg_out = None   # rows per frame on each side when several frames share a pass


class SparseConvTensor:
    """features [N,C] f32, indices [N,4] i32 (b,z,y,x), spatial_shape (D,H,W)."""

    def __init__(self, features, indices, spatial_shape, batch_size, table=None):
        self.features = features
        self.indices = indices
        self.spatial_shape = [int(s) for s in spatial_shape]
        self.batch_size = batch_size
        self.indice_dict = {}
        self.table = table              # kernels.SiteTable or None (built on demand)
        self.n = int(features.shape[0])
        self.seg = None                 # rows per frame when several frames share this tensor
        self.frames = None              # per-frame (row offset, rows, SiteTable) of this level, for dense()

    def _site_table(self):
        if self.table is None:
            self.table = K.site_table_build(self.indices, self.n, self.spatial_shape)
        return self.table

    def find_indice_pair(self, key):
        return self.indice_dict.get(key) if key is not None else None

    def shadow(self, features, indices=None, spatial_shape=None, table=None, n=None, seg=None, frames=None,
               new_sites=False):
        keep = indices is None and not new_sites
        t = SparseConvTensor(features, self.indices if keep else indices,
                             self.spatial_shape if spatial_shape is None else spatial_shape,
                             self.batch_size, self.table if keep else table)
        t.indice_dict = self.indice_dict
        if n is not None:
            t.n = n
        t.seg = self.seg if keep else seg
        t.frames = self.frames if keep else frames
        return t

    def dense_frames(self):
        """One dense [1, C*D, H, W] map per frame of a multi-frame tensor (middle.py:240-243 per frame)."""
        assert self.frames is not None
        D, H, W = self.spatial_shape
        outs = _DenseFramesFn.apply(self.features, self)
        return [o.view(1, self.features.shape[1] * D, H, W) for o in outs]

    def dense(self):
        """[B, C, D, H, W] (`SparseConvTensor.dense()`, used at middle.py:240)."""
        assert self.batch_size == 1
        D, H, W = self.spatial_shape
        out = _DenseFn.apply(self.features, self, self.features.shape[1])
        return out.view(1, self.features.shape[1], D, H, W)


# Layers with Cin, Cout in {32, 64} run on the tcgen05 tensor-core kernel (csrc/spconv_tc.cu, split-TF32,
# FP32-level accuracy); the narrow layers (7/16 channels) stay on the FP32 FFMA kernel.
# RSLO_SPCONV_TC=0 forces the FFMA kernel everywhere.
USE_TC = os.environ.get("RSLO_SPCONV_TC", "1") != "0"
USE_OWN_BN1D = os.environ.get("RSLO_OWN_BN1D", "1") != "0"     # A/B switch: 0 = torch native_batch_norm per frame


class _DenseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, st, C_):
        ctx.st = st
        ctx.C = C_
        return K.dense_from_sites(feat.contiguous(), st._site_table())

    @staticmethod
    def backward(ctx, g):
        st = ctx.st
        return K.dense_backward(g, st.indices, st.n, st.spatial_shape, ctx.C), None, None


class _DenseFramesFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, st):
        ctx.st = st
        feat = feat.contiguous()
        outs = []
        for off, rows, table, _ in st.frames:
            outs.append(K.dense_from_sites(feat[off:off + rows], table))
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gs):
        st = ctx.st
        C_ = gs[0].shape[0] // st.spatial_shape[0]
        grad = torch.empty((st.n, C_), dtype=gs[0].dtype, device=gs[0].device)
        for g, (off, rows, table, coors) in zip(gs, st.frames):
            if rows > 0:
                grad[off:off + rows] = K.dense_backward(g, coors, rows, st.spatial_shape, C_)
        return grad, None


_IMAGE_GENERATION = [0]


def invalidate_weight_images():
    """Forget every cached weight image.  In-place writes through `.data` / raw pointers (the reference's optimizer
    wrapper `p.data.mul_`, `model.data.copy_(master)`, NCCL broadcasts, this repo's fused Adam kernel) do not advance
    a tensor's version counter, so the network calls this at the start of every training forward (weights change
    every step anyway: an image is never reused across steps) and after load_state_dict / optimizer steps."""
    _IMAGE_GENERATION[0] += 1


class _ImageCache:
    """Pre-swizzled split-TF32 weight images of one layer (csrc/spconv_tc.cu), rebuilt when the weight changes
    (tensor version counter, storage address, or `invalidate_weight_images()`): the frames / pairs of a step and
    consecutive eval calls share them."""

    def __init__(self):
        self._c = {}

    def get(self, weight, transpose, mirror):
        key = (bool(transpose), bool(mirror))
        ver = (weight._version, weight.data_ptr(), _IMAGE_GENERATION[0])
        hit = self._c.get(key)
        if hit is None or hit[0] != ver:
            hit = (ver, K.spconv_tc_prepare(weight.detach(), transpose=transpose, mirror=mirror))
            self._c[key] = hit
        return hit[1]


class _SpConvFn(torch.autograd.Function):
    """out = act(bias + sum_k in[nbr[:,k]] @ W[k]); LeakyReLU fused (slope > 0 keeps it invertible
    from the saved output's sign)."""

    @staticmethod
    def forward(ctx, feat, weight, bias, entry, inverse, act, slope, images=None):
        nbr = entry.nbr_t if inverse else entry.nbr
        n_out = entry.n_in if inverse else entry.n_out
        feat = feat.contiguous()
        Kk, Cin, Cout = weight.shape
        # 16-channel layers: the tensor-core forward works (tested) but is slower than the FFMA kernel there
        # (half of each 32-wide K step is padding; measured 34.8 vs 32.2 ms/step), so only >= 32 channels use it
        ctx.tc = USE_TC and min(Cin, Cout) >= 32 and K.spconv_tc_supported(Cin, Cout, Kk)
        if ctx.tc:
            ctx.images = images if images is not None else _ImageCache()
            img = ctx.images.get(weight, False, False)
            out = K.spconv_tc_forward(feat, nbr, n_out, img, Cin, Cout, bias, act=act, slope=slope)
        else:
            out = K.spconv_forward(feat, nbr, n_out, weight, bias, act=act, slope=slope)
        ctx.entry, ctx.inverse, ctx.act, ctx.slope = entry, inverse, act, slope
        ctx.has_bias = bias is not None
        ctx.save_for_backward(feat, weight, out if act else None)
        return out

    @staticmethod
    def backward(ctx, g):
        feat, weight, out = ctx.saved_tensors
        e, inverse = ctx.entry, ctx.inverse
        # LeakyReLU backward (from the saved output) and the bias gradient in one kernel
        need_gb = ctx.has_bias and ctx.needs_input_grad[2]
        g, gb_fused = K.act_backward(g, out, ctx.act, ctx.slope, need_bias=need_gb)
        nbr = e.nbr_t if inverse else e.nbr
        nbr_t = e.nbr if inverse else e.nbr_t
        n_out = e.n_in if inverse else e.n_out
        n_in = e.n_out if inverse else e.n_in
        gi = gw = gb = None
        if ctx.needs_input_grad[0]:
            if ctx.tc:
                Kk, Cin, Cout = weight.shape
                img_t = ctx.images.get(weight, True, e.mirror)
                gi = K.spconv_tc_forward(g, nbr_t, n_in, img_t, Cout, Cin)
            else:
                gi = K.spconv_backward_data(g, nbr_t, n_in, weight, e.mirror)
        if ctx.needs_input_grad[1]:
            if USE_TC and K.spconv_tc_wgrad_supported(weight.shape[-2], weight.shape[-1]):
                gw = K.spconv_tc_backward_weight(feat, g, nbr, n_out, weight.shape)
            else:
                gw, _ = K.spconv_backward_weight(feat, g, nbr, n_out, weight.shape, need_bias=False)
        return gi, gw, gb_fused, None, None, None, None, None


class _SegBNActFn(torch.autograd.Function):
    """BatchNorm1d (+ LeakyReLU) over the rows of stacked frames with per-frame batch statistics (csrc/bn1d_seg.cu);
    updates the module's running statistics frame after frame like the reference's per-frame encoder calls."""

    @staticmethod
    def forward(ctx, feat, gamma, beta, mod, seg, slope):
        feat = feat.contiguous()
        training = mod.training
        z, mean_rstd = K.bn1d_seg_forward(feat, seg, gamma, beta, mod.running_mean, mod.running_var,
                                          mod.num_batches_tracked, mod.eps, mod.momentum, training, slope)
        ctx.seg, ctx.slope, ctx.training = seg, slope, training
        ctx.save_for_backward(feat, gamma, beta, mean_rstd)
        return z

    @staticmethod
    def backward(ctx, dz):
        feat, gamma, beta, mean_rstd = ctx.saved_tensors
        dx, dgamma, dbeta = K.bn1d_seg_backward(dz.contiguous(), feat, ctx.seg, mean_rstd, gamma, beta, ctx.slope,
                                                ctx.training)
        return dx, dgamma, dbeta, None, None, None


def _triple(v):
    return tuple(v) if isinstance(v, (list, tuple)) else (v, v, v)


class SparseConvolution(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=0, bias=True, subm=False,
                 inverse=False, indice_key=None):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.padding = _triple(kernel_size), _triple(stride), _triple(padding)
        self.subm, self.inverse, self.indice_key = subm, inverse, indice_key
        self.weight = nn.Parameter(torch.empty(*self.kernel_size, in_channels, out_channels))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        self._images = _ImageCache()
        self.fused_act = 0          # set by SparseSequential when a LeakyReLU follows
        self.fused_slope = 0.01
        self.reset_parameters()

    def reset_parameters(self):
        # spconv 1.x: kaiming_uniform_(a=sqrt(5)) on the [k,k,k,Cin,Cout] tensor, bias U(+-1/sqrt(fan_in))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in, _ = nn.init._calculate_fan_in_and_fan_out(self.weight)
            bound = 1 / math.sqrt(fan_in)
            nn.init.uniform_(self.bias, -bound, bound)

    def _entry(self, x):
        e = x.find_indice_pair(self.indice_key)
        if self.inverse:
            assert e is not None and e.kind == "strided", "inverse conv needs the keyed strided conv's tables"
            return e
        if e is not None:
            return e
        if self.subm:
            nbr = K.subm_table(x.indices, x.n, x._site_table(), self.kernel_size)
            e = IndexEntry("subm", nbr, nbr, x.n, x.n, x.indices, x.spatial_shape, x.table, True)
        else:
            tab, oc, n_out_dev, nbr, nbr_inv = K.strided_table(x.indices, x.n, x.spatial_shape, self.kernel_size,
                                                              self.stride, self.padding)
            n_out = int(n_out_dev.item())           # host sync: only on the on-demand path
            e = IndexEntry("strided", nbr, nbr_inv, x.n, n_out, oc, list(tab.shape), tab, False)
        if self.indice_key is not None:
            x.indice_dict[self.indice_key] = e
        return e

    def forward(self, x):
        e = self._entry(x)
        w = self.weight.view(-1, self.in_channels, self.out_channels)
        out = _SpConvFn.apply(x.features, w, self.bias, e, self.inverse, self.fused_act, self.fused_slope,
                              self._images)
        if self.inverse:
            # lands on the INPUT site set of the keyed conv
            return x.shadow(out, e.in_indices, e.in_shape, e.in_table, e.n_in, seg=e.seg_in,
                            frames=getattr(e, "in_frames", None), new_sites=True)
        if self.subm:
            return x.shadow(out)
        return x.shadow(out, e.out_indices, e.out_shape, e.out_table, e.n_out, seg=e.seg_out,
                        frames=getattr(e, "out_frames", None), new_sites=True)


class SubMConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True, indice_key=None):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, bias, subm=True,
                         indice_key=indice_key)


class SparseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True, indice_key=None):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, bias, indice_key=indice_key)

    def _entry(self, x):
        e = super()._entry(x)
        if not hasattr(e, "in_indices"):
            e.in_indices, e.in_shape, e.in_table = x.indices, x.spatial_shape, x.table
        return e


class SparseInverseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, indice_key, bias=True):
        super().__init__(in_channels, out_channels, kernel_size, bias=bias, inverse=True, indice_key=indice_key)


class SparseSequential(nn.Sequential):
    """Applies sparse modules to the SparseConvTensor and dense modules to `.features`
    (spconv.SparseSequential).  A LeakyReLU that directly follows a sparse conv (possibly through an
    `Empty` norm) is fused into the conv kernel's epilogue."""

    def __init__(self, *args):
        super().__init__(*args)
        self._plan = None

    def _build_plan(self):
        mods = list(self._modules.values())
        plan, i = [], 0
        while i < len(mods):
            m = mods[i]
            if (isinstance(m, nn.BatchNorm1d) and m.affine and m.track_running_stats and m.momentum is not None
                    and m.num_features % 4 == 0 and USE_OWN_BN1D):
                # BatchNorm1d (+ LeakyReLU) over the stacked frames: csrc/bn1d_seg.cu, per-frame statistics
                if i + 1 < len(mods) and isinstance(mods[i + 1], nn.LeakyReLU):
                    plan.append((m, "bn", mods[i + 1].negative_slope))
                    i += 2
                else:
                    plan.append((m, "bn", -1.0))
                    i += 1
                continue
            if isinstance(m, SparseConvolution):
                j = i + 1
                while j < len(mods) and type(mods[j]).__name__ == "Empty":
                    j += 1
                if j < len(mods) and isinstance(mods[j], nn.LeakyReLU):
                    plan.append((m, 1, mods[j].negative_slope))
                    i = j + 1
                    continue
                plan.append((m, 0, 0.0))
            else:
                plan.append((m, None, None))
            i += 1
        return plan

    def forward(self, x):
        if self._plan is None:
            self._plan = self._build_plan()
        for m, act, slope in self._plan:
            if act == "bn":
                if isinstance(x, SparseConvTensor) and x.features.is_cuda:
                    if x.n > 0:
                        seg = tuple(x.seg) if x.seg is not None else (x.features.shape[0],)
                        x = x.shadow(_SegBNActFn.apply(x.features, m.weight, m.bias, m, seg, slope))
                else:                               # dense / CPU input: the torch modules
                    f = x.features if isinstance(x, SparseConvTensor) else x
                    f = m(f)
                    f = F.leaky_relu(f, slope) if slope >= 0 else f
                    x = x.shadow(f) if isinstance(x, SparseConvTensor) else f
            elif isinstance(m, SparseConvolution):
                m.fused_act, m.fused_slope = act, slope
                x = m(x)
            elif isinstance(x, SparseConvTensor):
                if x.n > 0:
                    if (x.seg is not None and len(x.seg) > 1 and isinstance(m, nn.modules.batchnorm._BatchNorm)
                            and m.training):
                        # the reference normalises every frame with its own batch statistics (one encoder
                        # call per frame) and updates the running statistics frame after frame
                        x = x.shadow(torch.cat([m(f) for f in torch.split(x.features, x.seg)], dim=0))
                    else:
                        x = x.shadow(m(x.features))
            else:
                x = m(x)
        return x
