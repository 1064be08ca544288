"""2-D convolution of the dense head on the tcgen05 sparse-convolution kernels.

A dense BEV map is a fully occupied "sparse" level: in NHWC (channels_last) memory a feature map IS the
row matrix [B*H*W, C] the sparse kernels work on, and a 3x3 / 1x1 convolution is a gather-GEMM over a static
neighbour table (-1 outside the padding border).  Forward and data gradient run on csrc/spconv_tc.cu
(split-TF32 tensor-core implicit GEMM, FP32-level accuracy: measured ~1e-6 relative, at or below cuDNN's own
FP32 algorithms which reach 2e-5 with Winograd/FFT); the weight gradient is a plain library GEMM-shaped
reduction and stays on cuDNN (`aten::convolution_backward`).  No layout conversion happens between layers:
activations stay channels_last, which torch's BatchNorm / ReLU / cat / upsample handle natively.

Replaces `nn.Conv2d` as used by `rslo/models/odom_pred_base.py:155-276`, `rslo/layers/MaskConv.py:20-73`;
parameters keep the reference's names and OIHW shapes.
"""
import os

import numpy as np
import torch
from torch import nn

from .. import kernels as K

# Opt-in (RSLO_HEAD_TC=1).  Round-1 measurement on B200: numerically fine (<= 5e-6 vs float64) but NOT faster
# than cuDNN's FP32 algorithms at the head's small maps (train step 33.9 vs 32.9 ms): the gather-GEMM kernel
# is paced by its per-step staging latency, not by the tensor pipe (32 % active).  Default: cuDNN FP32.
USE_TC = os.environ.get("RSLO_HEAD_TC", "0") == "1"

_TABLES = {}


def grid_tables(B, H, W, ksize, stride, pad, device):
    """Static neighbour tables of a [B,H,W] grid for a ksize x ksize / stride / pad convolution.
    -> (nbr [B*Ho*Wo, K] input row per offset or -1, nbr_t, mirror, Ho, Wo): the data gradient is the gather
    conv over nbr_t with the transposed filters (mirrored offsets when stride == 1, where nbr_t is nbr)."""
    key = (B, H, W, ksize, stride, pad, str(device))
    hit = _TABLES.get(key)
    if hit is not None:
        return hit
    Ho, Wo = (H + 2 * pad - ksize) // stride + 1, (W + 2 * pad - ksize) // stride + 1
    ky, kx = np.meshgrid(np.arange(ksize), np.arange(ksize), indexing="ij")
    ky, kx = ky.reshape(-1), kx.reshape(-1)
    b, yo, xo = np.meshgrid(np.arange(B), np.arange(Ho), np.arange(Wo), indexing="ij")
    yi = yo.reshape(-1, 1) * stride + ky[None] - pad
    xi = xo.reshape(-1, 1) * stride + kx[None] - pad
    ok = (yi >= 0) & (yi < H) & (xi >= 0) & (xi < W)
    nbr = np.where(ok, (b.reshape(-1, 1) * H + yi) * W + xi, -1).astype(np.int32)
    if stride == 1 and Ho == H and Wo == W:
        nbr_t, mirror = nbr, True
    else:
        b, y, x = np.meshgrid(np.arange(B), np.arange(H), np.arange(W), indexing="ij")
        ny = y.reshape(-1, 1) + pad - ky[None]
        nx = x.reshape(-1, 1) + pad - kx[None]
        ok = (ny >= 0) & (nx >= 0) & (ny % stride == 0) & (nx % stride == 0)
        oy, ox = ny // stride, nx // stride
        ok &= (oy < Ho) & (ox < Wo)
        nbr_t, mirror = np.where(ok, (b.reshape(-1, 1) * Ho + oy) * Wo + ox, -1).astype(np.int32), False
    nbr_d = torch.from_numpy(np.ascontiguousarray(nbr)).to(device)
    nbr_t_d = nbr_d if nbr_t is nbr else torch.from_numpy(np.ascontiguousarray(nbr_t)).to(device)
    hit = (nbr_d, nbr_t_d, mirror, Ho, Wo)
    _TABLES[key] = hit
    return hit


def _rows(x):
    """[B,C,H,W] (any strides) -> ([B*H*W, C] row matrix sharing channels_last storage, channels_last tensor)."""
    x = x.contiguous(memory_format=torch.channels_last)
    B, C, H, W = x.shape
    return x.permute(0, 2, 3, 1).reshape(B * H * W, C), x


class _Conv2dTCFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, stride, pad):
        B, Cin, H, W = x.shape
        Cout, _, ks, _ = weight.shape
        nbr, nbr_t, mirror, Ho, Wo = grid_tables(B, H, W, ks, stride, pad, x.device)
        rows, x_cl = _rows(x)
        w9 = weight.permute(2, 3, 1, 0).reshape(ks * ks, Cin, Cout).contiguous()
        img = K.spconv_tc_prepare(w9)
        out = K.spconv_tc_forward(rows, nbr, B * Ho * Wo, img, Cin, Cout, bias)
        ctx.save_for_backward(x_cl, weight)
        ctx.geom = (nbr_t, mirror, stride, pad, bias is not None, (B, H, W, Ho, Wo))
        return out.view(B, Ho, Wo, Cout).permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, g):
        x_cl, weight = ctx.saved_tensors
        nbr_t, mirror, stride, pad, has_bias, (B, H, W, Ho, Wo) = ctx.geom
        Cout, Cin, ks, _ = weight.shape
        g_rows, g_cl = _rows(g)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            w9 = weight.permute(2, 3, 1, 0).reshape(ks * ks, Cin, Cout).contiguous()
            img_t = K.spconv_tc_prepare(w9, transpose=True, mirror=mirror)
            gi = K.spconv_tc_forward(g_rows, nbr_t, B * H * W, img_t, Cout, Cin)
            gx = gi.view(B, H, W, Cin).permute(0, 3, 1, 2)
        if ctx.needs_input_grad[1]:
            gw = torch.ops.aten.convolution_backward(g_cl, x_cl, weight, None, [stride, stride], [pad, pad], [1, 1],
                                                     False, [0, 0], 1, [False, True, False])[1]
        if has_bias and ctx.needs_input_grad[2]:
            gb = g_rows.sum(dim=0)
        return gx, gw, gb, None, None


class Conv2dTC(nn.Conv2d):
    """`nn.Conv2d` whose forward / data gradient run on the tensor-core gather-GEMM kernels when the shape
    allows (square kernel, symmetric padding, groups 1, Cin and Cout in {32,64,128,192,256,512}); the few
    narrow output convolutions (7- and 1-channel 1x1 heads) stay on cuDNN."""

    def _conv_forward(self, input, weight, bias):
        ks, st, pd = self.kernel_size, self.stride, self.padding
        if (USE_TC and input.is_cuda and self.groups == 1 and self.dilation == (1, 1) and ks[0] == ks[1]
                and st[0] == st[1] and isinstance(pd, tuple) and pd[0] == pd[1] and self.padding_mode == "zeros"
                and K.spconv_tc_supported(self.in_channels, self.out_channels, ks[0] * ks[1])):
            return _Conv2dTCFn.apply(input, weight, bias, st[0], pd[0])
        return super()._conv_forward(input, weight, bias)
