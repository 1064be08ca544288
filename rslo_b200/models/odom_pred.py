"""Odometry head (`rslo/models/odom_pred.py:44-435`, `odom_pred_base.py:25-346`,
`custom_resnet_spc.py:224-298`): masked ResNet encoder-decoder over the concatenated BEV maps of a
frame pair -> dense per-cell (t,q) map + softmax confidences -> confidence-voted global (t,q).

Module tree and parameter names follow the reference so its checkpoints load unchanged (485 state_dict
entries, tests/golden/state_dict_shapes.json).  Differences that do not change any output:
  * MaskConv mask propagation through the encoder is skipped (the propagated masks are never read);
  * the T=1 and T=20 confidences share one pass of the confidence conv stack (the reference runs the
    stack twice on the same values, `odom_pred.py:242-258`);
  * cell-anchor grids are cached instead of rebuilt per call.
The trunk (every convolution, BatchNorm, ReLU, residual add, upsample + concat) runs on the repo's own kernels:
layers/head_tc.py over csrc/conv2d_tc.cu (tcgen05 split-TF32 implicit GEMM) and csrc/head_ops.cu.
"""
import os

import torch
from torch import nn
from torch.nn import functional as F

from ..data.dataset import from_pointwise_local_transformation_tch
from ..layers.common import ParameterLayer
from ..layers.confidence import ConfidenceModule
from ..layers.conv2d_tc import Conv2dTC
from ..layers.head_tc import HeadTrunkEngine, head_trunk
from ..layers.pose_tail import head_geometry, head_tail
from ..layers.MaskConv import MaskConv
from ..layers.SparseConv import SPC_BN2d, SPC_ReLU, SPC_SyncBN2d
from ..torchplus import Empty, change_default_args
from ..utils.pose_utils import rotate_vec_by_q

REGISTERED_ODOM_PRED_CLASSES = {}

# The 2-D convolutions run in true FP32: cuDNN's default TF32 path gives ~1e-3 pose error, outside the
# 1e-4 relative parity bound of the path (measured on B200, tests/test_gpu_pair.py).
HEAD_ALLOW_TF32 = False
# The trunk runs on the repo's own tcgen05 convolutions + fused BN/ReLU kernels (layers/head_tc.py).
# RSLO_HEAD_TC=0 switches to torch/cuDNN FP32 for A/B comparison only.
USE_OWN_TRUNK = os.environ.get("RSLO_HEAD_TC", "1") != "0"
# The tail after the last convolution (confidences, local->global, vote, pyramid masks) is one fused kernel
# (layers/pose_tail.py); RSLO_HEAD_TAIL=0 runs the same steps as torch ops (A/B comparison).
USE_OWN_TAIL = os.environ.get("RSLO_HEAD_TAIL", "1") != "0"
# cuDNN autotuning of the head's FP32 convolutions (shapes are static; tuned once before graph capture).
HEAD_CUDNN_BENCHMARK = os.environ.get("RSLO_CUDNN_BENCHMARK", "1") != "0"


def register_odom_pred(cls, name=None):
    name = cls.__name__ if name is None else name
    assert name not in REGISTERED_ODOM_PRED_CLASSES, f"exist class: {REGISTERED_ODOM_PRED_CLASSES}"
    REGISTERED_ODOM_PRED_CLASSES[name] = cls
    return cls


def get_odom_class(name):
    assert name in REGISTERED_ODOM_PRED_CLASSES, f"available class: {REGISTERED_ODOM_PRED_CLASSES}"
    return REGISTERED_ODOM_PRED_CLASSES[name]


def SPC_add(a, b):
    if isinstance(a, (list, tuple)):
        m = None if a[1] is None or b[1] is None else ((a[1] + b[1]) / 2).float()
        return [a[0] + b[0], m]
    return a + b


class BasicBlock(nn.Module):
    """`custom_resnet_spc.py:224-298` with use_se = use_sa = False."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, BN=None, Conv2d=None, groups=1):
        super().__init__()
        self.conv1 = Conv2d(inplanes, planes, kernel_size=3, stride=stride, padding=1, bias=False, groups=groups)
        self.bn1 = BN(planes)
        self.relu = SPC_ReLU(inplace=True)
        self.conv2 = Conv2d(planes, planes, kernel_size=3, stride=1, padding=1, bias=False, groups=groups)
        self.bn2 = BN(planes)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        residual = x
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        if self.downsample is not None:
            residual = self.downsample(x)
        return self.relu(SPC_add(out, residual))


def _conf_stack(cin, BN, ReLU):
    return nn.Sequential(Conv2dTC(cin, 64, kernel_size=3, padding=1), BN(64), ReLU(),
                         Conv2dTC(64, 32, kernel_size=3, padding=1), BN(32), ReLU(),
                         Conv2dTC(32, 1, kernel_size=1))


class _FlatHead(nn.Module):
    """Tensor-in / tensor-out view of the head for CUDA-graph capture (shares the head's parameters;
    never registered as a child, so state_dict keys are unchanged)."""

    def __init__(self, head, n_levels, imgs_per_group):
        super().__init__()
        self.head = head
        self.n_levels = n_levels
        self.imgs_per_group = imgs_per_group

    def forward(self, *bevs):
        d = self.head._forward(list(bevs), imgs_per_group=self.imgs_per_group)
        out = [d["translation_preds"][0], d["rotation_preds"][0], d["tq_map_g"], d["t_conf"], d["r_conf"]]
        for pred, mask in d["pyramid_motion"]:
            out += [pred, mask]
        return tuple(out)


@register_odom_pred
class UNRResNetOdomPredEncDecSVDTempMask(nn.Module):
    def __init__(self, point_cloud_range=None, enc_use_norm=True, seq_len=1, layer_nums=(3, 5, 5),
                 layer_strides=(2, 2, 2), num_filters=(128, 128, 256), upsample_strides=(1, 2, 4),
                 num_upsample_filters=(256, 256, 256), num_input_features=128, encode_background_as_zeros=True,
                 use_groupnorm=False, bn_type="BN", num_groups=32, dropout=0.2, pooling_type="avg_pool",
                 pooling_size=1, cycle_constraint=False, conv_type="official", odom_format="rx+t",
                 pred_pyramid_motion=False, use_deep_supervision=False, use_loss_mask=True,
                 use_dynamic_mask=False, dense_predict=False, use_correlation=False, conf_type="linear",
                 use_SPGN=False, sync_bn=False, use_leakyReLU=False, dropout_input=False, first_conv_groups=1,
                 use_se=False, use_sa=False, use_svd=False, cubic_pred_height=0, name="odomPred", **kwargs):
        super().__init__()
        assert conv_type == "mask_conv", "only conv_type 'mask_conv' is built (shipped configs)"
        assert odom_format in ["rx+t", "r(x+t)"]
        assert bn_type in ["BN", "SyncBN"], "only (Sync)BN heads are built (shipped configs)"
        assert conf_type in ["linear", "softmax"]
        assert not (use_groupnorm or use_dynamic_mask or use_correlation or use_SPGN or use_leakyReLU or
                    dropout_input or use_se or use_sa or use_svd), "option outside the shipped configs"
        assert dropout > 0
        layer_nums, layer_strides = list(layer_nums), list(layer_strides)
        num_filters, upsample_strides = list(num_filters), list(upsample_strides)
        num_upsample_filters = list(num_upsample_filters)
        self.name = name
        self.conf_type = conf_type
        self._cubic_pred_height = cubic_pred_height
        self.point_cloud_range = point_cloud_range
        self.odom_format = odom_format
        self._use_mask_conv = True
        self._use_sparse_conv = False
        self.dense_predict = dense_predict
        self.use_svd = use_svd
        self._cycle_constraint = cycle_constraint
        self._num_input_features = num_input_features
        self._enc_use_norm = enc_use_norm
        self.pred_pyramid_motion = use_deep_supervision          # (sic) odom_pred_base.py:112

        bn_cls = SPC_SyncBN2d if (bn_type == "SyncBN" or sync_bn) else SPC_BN2d
        self.BatchNorm2d = change_default_args(eps=1e-3, momentum=0.01)(bn_cls)
        self.ReLU = SPC_ReLU
        Conv2d = change_default_args(bias=True)(Conv2dTC)
        BN, ReLU = self.BatchNorm2d, self.ReLU

        in_filters = [num_input_features, *num_filters[:-1]]
        blocks, skip_blocks, deblocks = [], [], []
        for i, layer_num in enumerate(layer_nums):
            block, nout = self._make_layer(in_filters[i], num_filters[i], layer_num, stride=layer_strides[i],
                                           first_groups=first_conv_groups if i == 0 else 1, use_norm=enc_use_norm)
            blocks.append(block)
            skip_blocks.append(nn.Sequential(Conv2d(nout, nout, kernel_size=3, stride=1, padding=1), BN(nout), ReLU()))
        py_blocks = []
        for i in range(len(num_upsample_filters)):
            cin = num_filters[-1] * 2 if i == 0 else num_upsample_filters[i - 1] + num_filters[-(i + 1)]
            deblocks.append(nn.Sequential(nn.Upsample(scale_factor=upsample_strides[i]),
                                          Conv2dTC(cin, num_upsample_filters[i], kernel_size=3, stride=1, padding=1),
                                          BN(num_upsample_filters[i]), ReLU()))
            if self.pred_pyramid_motion:
                c = num_upsample_filters[i]
                py_blocks.append(nn.Sequential(Conv2dTC(c, c // 2, kernel_size=3, stride=1, padding=1), BN(c // 2), ReLU(),
                                               Conv2dTC(c // 2, 64, kernel_size=3, stride=1, padding=1), BN(64), ReLU(),
                                               Conv2dTC(64, 7, 1, stride=1)))
        if self.pred_pyramid_motion:
            self.mask_gen_pools = nn.ModuleList([nn.MaxPool2d(kernel_size=3, stride=s, padding=1) for s in upsample_strides])
        self.blocks = nn.ModuleList(blocks)
        self.deblocks = nn.ModuleList(deblocks)
        self.skip_blocks = nn.ModuleList(skip_blocks)
        self.pyramid_motion_blocks = nn.ModuleList(py_blocks)
        c_last = num_upsample_filters[-1]
        self.tq_map_conv = nn.Sequential(Conv2dTC(c_last, 64, kernel_size=3, padding=1), BN(64), ReLU(),
                                         Conv2dTC(64, 32, kernel_size=3, padding=1), BN(32), ReLU(),
                                         Conv2dTC(32, 7, kernel_size=1))
        self.q_map_conf = ConfidenceModule(_conf_stack(c_last, BN, ReLU), conf_type=conf_type)
        self.t_map_conf = ConfidenceModule(_conf_stack(c_last, BN, ReLU), conf_type=conf_type)
        self.pool = nn.AdaptiveAvgPool2d((pooling_size, pooling_size)) if pooling_type == "avg_pool" \
            else nn.AdaptiveMaxPool2d((pooling_size, pooling_size))
        self.fc1 = nn.Linear(num_filters[-1] * pooling_size * pooling_size * seq_len, 1024)
        self.odom_dropout = nn.Dropout(p=dropout)
        self.dense_dropout = nn.Dropout2d(p=dropout)
        self.fc2 = nn.Linear(1024, 7)
        self.softmax = nn.Softmax(dim=-1)
        self.SPGN = Empty()
        self.dynamic_sigma = ParameterLayer(torch.ones(1) * 0.1, requires_grad=True)
        self.pyramid_tconf_blocks = nn.ModuleList(
            [ConfidenceModule(_conf_stack(c, BN, ReLU), conf_type=conf_type) for c in num_upsample_filters]
            if self.pred_pyramid_motion else [])
        self.pyramid_qconf_blocks = nn.ModuleList(
            [ConfidenceModule(_conf_stack(c, BN, ReLU), conf_type=conf_type) for c in num_upsample_filters]
            if self.pred_pyramid_motion else [])
        self.hier_weight_gen = nn.AvgPool2d(3, 2, padding=1)

        # init (`odom_pred.py:381-389`)
        for m in self.modules():
            if isinstance(m, nn.Conv2d) and m.weight.requires_grad:
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d)):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def _make_layer(self, inplanes, planes, num_blocks, stride=1, first_groups=1, use_norm=True):
        conv2d = change_default_args(propagate_mask=False)(MaskConv)
        BN = self.BatchNorm2d if use_norm else Empty
        downsample = None
        if stride != 1 or inplanes != planes:
            downsample = nn.Sequential(conv2d(inplanes, planes, kernel_size=1, stride=stride, bias=False,
                                              groups=first_groups), BN(planes))
        layers = [BasicBlock(inplanes, planes, stride, downsample, BN=BN, Conv2d=conv2d, groups=first_groups)]
        for _ in range(1, num_blocks):
            layers.append(BasicBlock(planes, planes, BN=BN, Conv2d=conv2d))
        return nn.Sequential(*layers), planes

    @staticmethod
    def create_cycle_constraint_data(xs):
        """all ordered pairs i<j of the sequence (`odom_pred_base.py:305-324`)."""
        assert len(xs) >= 2
        b, C, H, W = xs[0].shape
        x1, x2 = [], []
        for i in range(len(xs)):
            for j in range(i + 1, len(xs)):
                x1.append(xs[i])
                x2.append(xs[j])
        return [torch.stack(x1, dim=1).reshape(-1, C, H, W), torch.stack(x2, dim=1).reshape(-1, C, H, W)]

    def forward(self, xs, tq_map_gt=None, local_spatial_features=None, **kwargs):
        """xs: the T frames' BEV maps, each [S,C,H,W] (S samples; the reference always has S = 1).  All ordered
        frame pairs i<j of every sample go through the head in one pass; BatchNorm batch statistics are per sample."""
        if not isinstance(xs, list):
            xs = [xs]
        ipg = len(xs) * (len(xs) - 1) // 2 if self._cycle_constraint else 1
        if self.use_cuda_graph and xs[0].is_cuda and self.dense_predict and self.pred_pyramid_motion:
            return self._forward_graphed(xs, ipg)
        return self._forward(xs, tq_map_gt, local_spatial_features, imgs_per_group=ipg, **kwargs)

    # ---- CUDA-graph replay of the head -----------------------------------------------------------
    # Every shape in the head is fixed by the BEV grid, so forward and backward are each captured
    # once per (frames, mode) into a CUDA graph (torch.cuda.make_graphed_callables) and replayed:
    # the ~450 kernel launches of a trunk pass and the tail's elementwise launches become two graph launches.
    use_cuda_graph = os.environ.get("RSLO_CUDA_GRAPHS", "1") != "0"

    def _apply(self, fn, *a, **k):
        # parameters may move (net.to(device)): captured graphs and prepared weight images point at the old storage
        self.__dict__.pop("_graphed", None)
        self.__dict__.pop("_mode_sig", None)
        e = self.__dict__.get("_trunk_engine")
        if e is not None:
            e.clear()
        return super()._apply(fn, *a, **k)

    def train(self, mode=True):
        self.__dict__.pop("_mode_sig", None)
        return super().train(mode)

    def _mode_signature(self):
        """what a captured graph bakes in besides shapes: per-layer BN mode (freeze_bn flips single layers to eval)
        and which parameters receive gradients (freeze_bn_affine).  Cached; `train()` / `eval()` on the head or any
        module above it (which is how the reference flips these, `voxel_odom_net.py:206-224`) and `_apply` reset it."""
        sig = self.__dict__.get("_mode_sig")
        if sig is None:
            bn = tuple(m.training for m in self.modules() if isinstance(m, nn.modules.batchnorm._BatchNorm))
            rg = tuple(p.requires_grad for p in self.parameters())
            sig = self.__dict__["_mode_sig"] = hash((bn, rg))
        return sig

    def _forward_graphed(self, xs, ipg):
        # one captured graph per stream: replays on different streams must not share static buffers
        key = (len(xs), self.training, self._mode_signature(), tuple(x.requires_grad for x in xs), tuple(xs[0].shape),
               xs[0].device.index, torch._C._cuda_getCurrentRawStream(xs[0].device.index), torch.is_grad_enabled())
        cache = self.__dict__.setdefault("_graphed", {})
        g = cache.get(key)
        if g is None:
            flat = _FlatHead(self, len(self.deblocks), ipg)
            flat.training = self.training
            bufs = {n: b.clone() for n, b in self.named_buffers()}
            sample = tuple(torch.randn_like(x).requires_grad_(x.requires_grad) for x in xs)
            from .._lib import lib
            # training graphs (forward + backward, ~880 nodes) are captured with programmatic dependent launch: +1 % on
            # the train step; the forward-only eval graph measured 13 % SLOWER with it (316 vs 362 pairs/s), so not there
            lib.rslo_set_graph_capture_hint(1 if self.training else 0)
            try:
                with torch.enable_grad():
                    g = torch.cuda.make_graphed_callables(flat, sample, allow_unused_input=True)
            finally:
                lib.rslo_set_graph_capture_hint(0)
            with torch.no_grad():                      # capture warm-up ran BN on the random sample
                for n, b in self.named_buffers():
                    b.copy_(bufs[n])
            cache[key] = g
        outs = g(*[x.contiguous() for x in xs])
        t, q, tq_map_g, t_conf, r_conf = outs[:5]
        pyramid = [[outs[5 + 2 * i], outs[6 + 2 * i]] for i in range((len(outs) - 5) // 2)]
        return {"translation_preds": [t], "rotation_preds": [q], "tq_map_g": tq_map_g, "pyramid_motion": pyramid,
                "transformed_inputs": None, "t_conf": t_conf, "r_conf": r_conf}

    # ---- trunk: BEV pair -> raw (t,q) map, confidence logits, pyramid predictions -------------------------
    def _engine(self):
        e = self.__dict__.get("_trunk_engine")
        if e is None:
            e = self.__dict__["_trunk_engine"] = HeadTrunkEngine(self)
        return e

    def _trunk_own(self, x1, x2, imgs_per_group):
        """csrc/conv2d_tc.cu + csrc/head_ops.cu (layers/head_tc.py): NHWC, split-TF32 tcgen05 convolutions."""
        outs = head_trunk(self._engine(), x1, x2, imgs_per_group or x1.shape[0])
        tq32, tl32, rl32, mask = outs[0], outs[1], outs[2], outs[-1]
        nchw = lambda t, c: t[..., :c].permute(0, 3, 1, 2)
        return nchw(tq32, 7), nchw(tl32, 1), nchw(rl32, 1), [nchw(p, 7) for p in outs[3:-1]], mask.unsqueeze(1)

    def _trunk_torch(self, x1, x2):
        """The same trunk through torch / cuDNN FP32 (A/B switch RSLO_HEAD_TC=0; never the default)."""
        input_mask = (torch.sum(x1, dim=1, keepdim=True) != 0).detach_().to(dtype=x1.dtype)
        x = torch.cat([x1, x2], dim=1)
        ups = []
        for i in range(len(self.blocks)):
            x = self.blocks[i](x)
            ups.append(self.skip_blocks[i](x[0]))
        x = x[0]
        py_raw = []
        for i in range(len(self.deblocks)):
            x = torch.cat([x, ups[-(i + 1)]], dim=1)
            x = self.deblocks[i](x)
            if self.pred_pyramid_motion and i < len(self.deblocks) - 1:
                py_raw.append(self.pyramid_motion_blocks[i](x))
        tq_map = self.tq_map_conv(x)
        t_logit = self.t_map_conf.conf_model(x)
        r_logit = self.q_map_conf.conf_model(x)
        if self.training:                   # second scoring pass of the reference (`odom_pred.py:242-258`): same
            with torch.no_grad():           # values, but the BN running statistics move a second time
                self.t_map_conf.conf_model(x)
                self.q_map_conf.conf_model(x)
        return tq_map, t_logit, r_logit, py_raw, input_mask

    def _forward(self, xs, tq_map_gt=None, local_spatial_features=None, imgs_per_group=None, **kwargs):
        if not isinstance(xs, list):
            xs = [xs]
        assert self.dense_predict, "only the dense-prediction head is built (shipped configs)"
        if self._cycle_constraint:
            xs = self.create_cycle_constraint_data(xs)
        x1, x2 = xs
        if USE_OWN_TRUNK and x1.is_cuda:
            if (USE_OWN_TAIL and self.pred_pyramid_motion and len(self.deblocks) == 3 and self.conf_type == "softmax"
                    and self.odom_format == "rx+t"):
                # trunk + fused tail kernel (csrc/pose_tail.cu): no torch op between the last convolution and the pose
                outs = head_trunk(self._engine(), x1, x2, imgs_per_group or x1.shape[0])
                geom = self.__dict__.get("_tail_geom")
                if geom is None or (geom.H, geom.W) != tuple(x1.shape[2:]):
                    geom = self.__dict__["_tail_geom"] = head_geometry(x1.shape[2], x1.shape[3], self.point_cloud_range)
                t, q, tq_map_g, t_conf, r_conf, pyramid = head_tail(outs[0], outs[1], outs[2], outs[3], outs[4], outs[5], geom)
                return {"translation_preds": [t], "rotation_preds": [q], "tq_map_g": tq_map_g, "pyramid_motion": pyramid,
                        "transformed_inputs": None, "t_conf": t_conf, "r_conf": r_conf}
            tq_map, t_logit, r_logit, py_raw, input_mask = self._trunk_own(x1, x2, imgs_per_group)
        else:
            with torch.backends.cudnn.flags(enabled=True, allow_tf32=HEAD_ALLOW_TF32, benchmark=HEAD_CUDNN_BENCHMARK):
                tq_map, t_logit, r_logit, py_raw, input_mask = self._trunk_torch(x1, x2)

        return self._tail_torch(tq_map, t_logit, r_logit, py_raw, input_mask)

    def _tail_torch(self, tq_map, t_logit, r_logit, py_raw, input_mask):
        """`odom_pred.py:210-313` after the convolutions as torch ops (A/B switch RSLO_HEAD_TAIL=0, CPU, and the
        float64 yardstick of tests/test_gpu_tail.py); the default GPU path is layers/pose_tail.py."""
        py_masks = []
        if self.pred_pyramid_motion:
            p_mask = input_mask
            for i in range(len(self.deblocks) - 1):
                p_mask = self.mask_gen_pools[-(i + 1)](p_mask)
                py_masks.append(p_mask)
            py_masks.reverse()
        py_preds = [[p * (m > 0).to(dtype=p.dtype), m] for p, m in zip(py_raw, py_masks)]
        q_map = tq_map[:, 3:] / torch.norm(tq_map[:, 3:], dim=1, keepdim=True)
        tq_map = torch.cat([tq_map[:, :3], q_map], dim=1)

        t_conf = self.t_map_conf(None, extra_mask=input_mask, logit=t_logit)
        r_conf = self.q_map_conf(None, extra_mask=input_mask, logit=r_logit)
        tq_map_g = from_pointwise_local_transformation_tch(tq_map, self.point_cloud_range)
        odoms = self.aggregate_tq([tq_map_g], t_confs=[t_conf], r_confs=[r_conf])
        temp_t_conf = self.t_map_conf(None, extra_mask=input_mask, temperature=20, logit=t_logit.detach())
        temp_r_conf = self.q_map_conf(None, extra_mask=input_mask, temperature=20, logit=r_logit.detach())
        temp_tq_conf = torch.cat([temp_t_conf, temp_r_conf], dim=1).detach()
        pyramid_motion = py_preds + [[tq_map * input_mask, input_mask * temp_tq_conf]]
        for p in range(2, len(pyramid_motion) + 1):
            pyramid_motion[-p][1] = pyramid_motion[-p][1] * self.hier_weight_gen(pyramid_motion[-(p - 1)][1])

        translations, rotations = [], []
        for x in odoms:
            translation, rotation = x[:, :3], x[:, 3:]
            if self.odom_format == "r(x+t)":
                translation = rotate_vec_by_q(translation, rotation)
            rotation = rotation / (torch.norm(rotation, dim=1, keepdim=True) + 1e-12)
            translations.append(translation)
            rotations.append(rotation)
        return {"translation_preds": translations, "rotation_preds": rotations, "tq_map_g": tq_map_g * input_mask,
                "pyramid_motion": pyramid_motion, "transformed_inputs": None, "t_conf": t_conf, "r_conf": r_conf}

    def aggregate_tq(self, tq_maps_g, t_confs, r_confs):
        """confidence-weighted vote (`odom_pred.py:347-357`, use_svd False)."""
        odoms = []
        for tq_map_g, t_conf, r_conf in zip(tq_maps_g, t_confs, r_confs):
            t = torch.sum(tq_map_g[:, :3] * t_conf, dim=(2, 3)) / (torch.sum(t_conf, dim=(2, 3)) + 1e-12)
            q = torch.sum(tq_map_g[:, 3:] * r_conf, dim=(2, 3)) / (torch.sum(r_conf, dim=(2, 3)) + 1e-12)
            odoms.append(torch.cat([t, q], dim=-1))
        return odoms
