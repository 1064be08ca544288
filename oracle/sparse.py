"""CPU restatement of the sparse 3-D encoder (a3-a7): VFE mean, SpMiddleFHDWithCov2_3.

TEST INFRASTRUCTURE ONLY.  The sparse-conv arithmetic belongs to the un-vendored spconv_plus fork
(parity unpinned); the layer list, activation and BN placement follow
`rslo/models/middle.py:119-245`, the VFE follows `rslo/models/voxel_encoder.py:272-280`.
Weights are read from a state_dict with the reference's key names
(`middle_feature_extractor.middle_conv.0.weight [kD,kH,kW,Cin,Cout]`, ...).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import native


def vfe_mean(voxels, num_points):
    """voxel_encoder.py:272-280."""
    voxels = torch.as_tensor(voxels)
    num_points = torch.as_tensor(num_points)
    mean = voxels[:, :, :7].sum(dim=1) / num_points.type_as(voxels).view(-1, 1)
    mean[:, 4:7] = mean[:, 4:7] / (torch.norm(mean[:, 4:7], dim=-1, keepdim=True) + 1e-12)
    return mean.contiguous()


def gather_conv(feat, nbr, weight, bias=None):
    """out[o] = bias + sum_k feat[nbr[o,k]] @ W[k]   (W [K,Cin,Cout], nbr [M,K] with -1 = none).

    Accumulates offset by offset like spconv's indice_conv (one GEMM per kernel offset)."""
    nbr = torch.as_tensor(nbr).long()
    M, K = nbr.shape
    W = weight.reshape(K, weight.shape[-2], weight.shape[-1])
    out = feat.new_zeros((M, W.shape[-1]))
    for k in range(K):
        sel = torch.nonzero(nbr[:, k] >= 0).squeeze(1)
        if sel.numel() == 0:
            continue
        out.index_add_(0, sel, feat[nbr[sel, k]] @ W[k])
    if bias is not None:
        out = out + bias
    return out


class Level:
    def __init__(self, coors, shape):
        self.coors = np.ascontiguousarray(coors, np.int32)   # [N,3] z,y,x
        self.shape = list(shape)


def build_tables(coors_zyx, sparse_shape):
    """All index tables of SpMiddleFHDWithCov2_3 for one frame (middle.py:119-213)."""
    L0 = Level(coors_zyx, sparse_shape)
    t = {"L0": L0}
    t["subm0"] = native.subm_table(L0.coors, L0.shape)
    c1, s1, t["conv3d2"], t["conv3d2_inv"] = native.strided_table(L0.coors, L0.shape, (3, 3, 3), (2, 2, 2), (1, 1, 1))
    L1 = Level(c1, s1)
    t["L1"] = L1
    t["subm1"] = native.subm_table(L1.coors, L1.shape)
    c2, s2, t["conv3d3"], t["conv3d3_inv"] = native.strided_table(L1.coors, L1.shape, (3, 3, 3), (2, 2, 2), (1, 1, 1))
    L2 = Level(c2, s2)
    t["L2"] = L2
    t["subm2"] = native.subm_table(L2.coors, L2.shape)
    c3, s3, t["conv3d4"], _ = native.strided_table(L2.coors, L2.shape, (3, 3, 3), (2, 2, 2), (0, 1, 1))
    L3 = Level(c3, s3)
    t["L3"] = L3
    t["subm3"] = native.subm_table(L3.coors, L3.shape)
    c4, s4, t["conv3d5"], _ = native.strided_table(L3.coors, L3.shape, (3, 1, 1), (2, 1, 1), (0, 0, 0))
    t["L4"] = Level(c4, s4)
    return t


def middle_forward(sd, voxel_features, coors_bzyx, sparse_shape, training=False,
                   prefix="middle_feature_extractor.", bn_stats_out=None):
    """SpMiddleFHDWithCov2_3.forward (middle.py:219-245) -> (bev [1,128,H,W], cov [N,7], tables)."""
    coors = np.asarray(coors_bzyx)[:, 1:4]
    t = build_tables(coors, sparse_shape)
    lrelu = lambda x: F.leaky_relu(x, 0.01)

    def conv(name, x, table):
        return gather_conv(x, t[table], sd[prefix + name + ".weight"], sd[prefix + name + ".bias"])

    x = torch.as_tensor(voxel_features)
    # middle_conv
    x = lrelu(conv("middle_conv.0", x, "subm0"))
    x = lrelu(conv("middle_conv.3", x, "subm0"))
    x = lrelu(conv("middle_conv.6", x, "conv3d2"))
    x = lrelu(conv("middle_conv.9", x, "subm1"))
    x = lrelu(conv("middle_conv.12", x, "subm1"))
    x = lrelu(conv("middle_conv.15", x, "conv3d3"))
    ret0 = x
    # middle_conv_tail
    x = lrelu(conv("middle_conv_tail.0", x, "subm2"))
    x = lrelu(conv("middle_conv_tail.3", x, "subm2"))
    x = lrelu(conv("middle_conv_tail.6", x, "subm2"))
    x = lrelu(conv("middle_conv_tail.9", x, "conv3d4"))
    x = lrelu(conv("middle_conv_tail.12", x, "subm3"))
    x = lrelu(conv("middle_conv_tail.15", x, "subm3"))
    x = lrelu(conv("middle_conv_tail.18", x, "subm3"))
    x = lrelu(conv("middle_conv_tail.21", x, "conv3d5"))
    # dense(): [1, C, D, H, W] -> view [1, C*D, H, W]   (middle.py:240-243)
    L4 = t["L4"]
    D, H, W = L4.shape
    dense = x.new_zeros((1, x.shape[1], D, H, W))
    c = torch.as_tensor(L4.coors).long()
    dense[0, :, c[:, 0], c[:, 1], c[:, 2]] = x.t()
    bev = dense.view(1, x.shape[1] * D, H, W)

    # middle_cov_deconv (real nn.BatchNorm1d eps 1e-5 momentum .1, middle.py:181-198)
    def bn(name, y):
        p = prefix + name
        if training:
            mean = y.mean(0)
            var = y.var(0, unbiased=False)
            if bn_stats_out is not None:
                bn_stats_out[p] = (mean, y.var(0, unbiased=True))
        else:
            mean, var = sd[p + ".running_mean"], sd[p + ".running_var"]
        return (y - mean) / torch.sqrt(var + 1e-5) * sd[p + ".weight"] + sd[p + ".bias"]

    y = ret0
    y = lrelu(bn("middle_cov_deconv.1", conv("middle_cov_deconv.0", y, "conv3d3_inv")))
    y = lrelu(bn("middle_cov_deconv.4", conv("middle_cov_deconv.3", y, "subm1")))
    y = lrelu(bn("middle_cov_deconv.7", conv("middle_cov_deconv.6", y, "conv3d2_inv")))
    y = lrelu(bn("middle_cov_deconv.10", conv("middle_cov_deconv.9", y, "subm0")))
    y = lrelu(bn("middle_cov_deconv.13", conv("middle_cov_deconv.12", y, "subm0")))
    y = conv("middle_cov_deconv.15", y, "subm0")
    y = torch.cat([F.elu(y[:, :3]) + 1 + 1e-6, y[:, 3:]], dim=1)      # middle.py:237
    return bev, y, t
