"""Sparse 3-D encoder with covariance decoder (`rslo/models/middle.py:37-245`)."""
import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

from .. import kernels as K
from ..layers import sparse3d as spconv
from ..layers.sparse3d import IndexEntry
from ..torchplus import Empty, change_default_args

REGISTERED_MIDDLE_CLASSES = {}


def register_middle(cls, name=None):
    name = cls.__name__ if name is None else name
    assert name not in REGISTERED_MIDDLE_CLASSES, f"exist class: {REGISTERED_MIDDLE_CLASSES}"
    REGISTERED_MIDDLE_CLASSES[name] = cls
    return cls


def get_middle_class(name):
    assert name in REGISTERED_MIDDLE_CLASSES, f"available class: {REGISTERED_MIDDLE_CLASSES}"
    return REGISTERED_MIDDLE_CLASSES[name]


def build_frame_tables(indices, n, sparse_shape, table0=None):
    """All index tables of SpMiddleFHDWithCov2_3 for one frame, enqueued without a host round trip
    (row counts of the deeper levels stay on the device while the next level is built) and finished
    by ONE device->host copy of the four counts.  Returns {indice_key: IndexEntry}."""
    geoms = [("conv3d2", (3, 3, 3), (2, 2, 2), (1, 1, 1)), ("conv3d3", (3, 3, 3), (2, 2, 2), (1, 1, 1)),
             ("conv3d4", (3, 3, 3), (2, 2, 2), (0, 1, 1)), ("conv3d5", (3, 1, 1), (2, 1, 1), (0, 0, 0))]
    subm_keys = ["subm0", "subm1", "subm2", "subm3"]
    caps = None
    while True:
        tab = table0 if table0 is not None else K.site_table_build(indices, n, sparse_shape)
        lv = [dict(idx=indices, cap=n, ndev=None, shape=list(sparse_shape), tab=tab)]
        raw = {}
        for li, (key, ks, st, pd) in enumerate(geoms):
            cur = lv[-1]
            raw[subm_keys[li]] = K.subm_table(cur["idx"], cur["cap"], cur["tab"], (3, 3, 3), n_dev=cur["ndev"])
            cap = None if caps is None else caps[li]
            if cap is None:
                # typical LiDAR levels shrink; overflow is detected below and retried exactly
                cells = int(np.prod(K.out_shape_of(cur["shape"], ks, st, pd)))
                cap = max(1, min(cur["cap"], cells))
            otab, oc, ndev2, nbr, nbr_inv = K.strided_table(cur["idx"], cur["cap"], cur["shape"], ks, st, pd,
                                                            n_dev=cur["ndev"], out_cap=cap)
            raw[key] = (nbr, nbr_inv)
            lv.append(dict(idx=oc, cap=cap, ndev=ndev2[:1], shape=list(otab.shape), tab=otab, ndev2=ndev2))
        counts = torch.stack([l["ndev2"] for l in lv[1:]]).cpu().numpy()    # the frame's only sync
        if (counts[:, 1] <= [l["cap"] for l in lv[1:]]).all():
            break
        caps = [int(c) for c in counts[:, 1]]
        caps = [max(int(c), 1) * 2 for c in counts[:, 1]]
    ns = [n] + [int(c) for c in counts[:, 0]]
    entries = {}
    for li in range(4):
        l = lv[li]
        nbr = raw[subm_keys[li]]
        entries[subm_keys[li]] = IndexEntry("subm", nbr, nbr, ns[li], ns[li], l["idx"], l["shape"], l["tab"], True)
    for li, (key, ks, st, pd) in enumerate(geoms):
        nbr, nbr_inv = raw[key]
        o = lv[li + 1]
        e = IndexEntry("strided", nbr, nbr_inv, ns[li], ns[li + 1], o["idx"], o["shape"], o["tab"], False)
        e.in_indices, e.in_shape, e.in_table = lv[li]["idx"], lv[li]["shape"], lv[li]["tab"]
        entries[key] = e
    # decoder keys are new names for site sets that already have tables (middle.py:181-213)
    entries["dsubm3"] = entries["subm1"]
    entries["dsubm2"] = entries["subm0"]
    entries["dsubm1"] = entries["subm0"]
    return entries


@register_middle
class SpMiddleFHDWithCov2_3(nn.Module):
    def __init__(self, output_shape, use_GN=False, sync_bn=False, bn_type="None", use_leakyReLU=False,
                 relu_type="ReLU", num_input_features=128, num_filters_down1=(64,), num_filters_down2=(64, 64),
                 name="SpMiddleFHDWithConf"):
        super().__init__()
        assert bn_type in ["None", "BN", "SyncBN", "SemiGlobalSyncBN", "MaskSyncBN"]
        assert relu_type in ["", "ReLU", "LeakyReLU", "PReLU"]
        self.name = name
        if bn_type != "None":
            # the shipped configs (bn_type "None") are the scope of this build
            raise NotImplementedError("SpMiddleFHDWithCov2_3: only bn_type 'None' is built")
        BatchNorm1d = Empty
        SpConv3d = change_default_args(bias=True)(spconv.SparseConv3d)
        SubMConv3d = change_default_args(bias=True)(spconv.SubMConv3d)
        ConvTranspose3d = change_default_args(bias=True)(spconv.SparseInverseConv3d)
        if use_leakyReLU or relu_type == "LeakyReLU":
            self.relu = nn.LeakyReLU
        else:
            raise NotImplementedError("SpMiddleFHDWithCov2_3: only LeakyReLU is built (shipped configs)")

        sparse_shape = np.array(output_shape[1:4]) + [1, 0, 0]
        self.sparse_shape = sparse_shape
        self.voxel_output_shape = output_shape

        self.middle_conv = spconv.SparseSequential(
            SubMConv3d(num_input_features, 16, 3, indice_key="subm0"), BatchNorm1d(16), self.relu(),
            SubMConv3d(16, 16, 3, indice_key="subm0"), BatchNorm1d(16), self.relu(),
            SpConv3d(16, 32, 3, 2, padding=1, indice_key="conv3d2"), BatchNorm1d(32), self.relu(),
            SubMConv3d(32, 32, 3, indice_key="subm1"), BatchNorm1d(32), self.relu(),
            SubMConv3d(32, 32, 3, indice_key="subm1"), BatchNorm1d(32), self.relu(),
            SpConv3d(32, 64, 3, 2, padding=1, indice_key="conv3d3"), BatchNorm1d(64), self.relu(),
        )
        self.middle_conv_tail = spconv.SparseSequential(
            SubMConv3d(64, 64, 3, indice_key="subm2"), BatchNorm1d(64), self.relu(),
            SubMConv3d(64, 64, 3, indice_key="subm2"), BatchNorm1d(64), self.relu(),
            SubMConv3d(64, 64, 3, indice_key="subm2"), BatchNorm1d(64), self.relu(),
            SpConv3d(64, 64, 3, 2, padding=[0, 1, 1], indice_key="conv3d4"), BatchNorm1d(64), self.relu(),
            SubMConv3d(64, 64, 3, indice_key="subm3"), BatchNorm1d(64), self.relu(),
            SubMConv3d(64, 64, 3, indice_key="subm3"), BatchNorm1d(64), self.relu(),
            SubMConv3d(64, 64, 3, indice_key="subm3"), BatchNorm1d(64), self.relu(),
            SpConv3d(64, 64, (3, 1, 1), (2, 1, 1), indice_key="conv3d5"), BatchNorm1d(64), self.relu(),
        )
        self.middle_cov_deconv = spconv.SparseSequential(
            ConvTranspose3d(64, 32, 3, indice_key="conv3d3"), nn.BatchNorm1d(32), self.relu(),
            SubMConv3d(32, 32, 3, indice_key="dsubm3"), nn.BatchNorm1d(32), self.relu(),
            ConvTranspose3d(32, 16, 3, indice_key="conv3d2"), nn.BatchNorm1d(16), self.relu(),
            SubMConv3d(16, 16, 3, indice_key="dsubm2"), nn.BatchNorm1d(16), self.relu(),
            SubMConv3d(16, 16, 3, indice_key="dsubm2"), nn.BatchNorm1d(16), self.relu(),
            SubMConv3d(16, 7, 3, indice_key="dsubm1"),
        )
        self.max_batch_size = 6

    def forward(self, voxel_features, coors, batch_size, table0=None):
        assert batch_size == 1, "Only support batch_size=1 for now"
        coors = coors.int().contiguous()
        n = int(voxel_features.shape[0])
        ret = spconv.SparseConvTensor(voxel_features, coors, self.sparse_shape, batch_size, table=table0)
        ret.indice_dict = build_frame_tables(coors, n, [int(s) for s in self.sparse_shape], table0)
        ret.table = ret.indice_dict["subm0"].out_table
        ret0 = self.middle_conv(ret)
        ret = self.middle_conv_tail(ret0)
        cov_pred = self.middle_cov_deconv(ret0)
        cov = cov_pred.features
        cov = torch.cat([F.elu(cov[:, :3]) + 1 + 1e-6, cov[:, 3:]], dim=1)     # middle.py:237
        ret = ret.dense()
        N, C, D, H, W = ret.shape
        ret = ret.view(N, C * D, H, W)
        return ret, cov
