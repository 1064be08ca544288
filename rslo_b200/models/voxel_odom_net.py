"""Network orchestrator `UnVoxelOdomNetICP3` (`rslo/models/voxel_odom_net.py:46-834`).

`net(example) -> dict` keeps the reference's contract.  Two input forms:
  * the reference's: ``example["voxels"|"num_points"|"coordinates"][t]`` produced by a voxel
    generator (CPU worker or `rslo_b200.builder.voxel_builder`), ``example["num_voxels"][t]``;
  * B200-native: ``example["points"][t]`` = raw scan ``[P,7]`` on the device; the fused
    scatter-voxeliser + VFE kernel (csrc/voxelize.cu) runs inside forward and also hands its site
    table to the sparse encoder, so the 11 MB/frame `voxels` tensor is never materialised.
Several independent samples of a step go through ONE call: ``example["n_samples"] = S`` with the S*T frames
listed sample-major.  Every sparse layer then launches once on all S*T frames, the head sees all S*pairs BEV pairs
in one pass (BatchNorm batch statistics stay per sample / per frame exactly as in S separate calls, running
statistics advance sample after sample) and ``loss`` is the mean over the samples, i.e. what S calls with gradient
accumulation of loss/S produce (the reference itself asserts batch_size == 1, `middle.py:221`).
"""
import time

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

from .. import kernels as K
from ..data.dataset import generate_pointwise_local_transformation_tch
from ..layers.pose_tail import loss_geometry, loss_tail
from ..layers.sparse3d import invalidate_weight_images
from ..torchplus import roll
from ..utils import pose_utils
from . import middle, odom_pred, voxel_encoder

REGISTERED_NETWORK_CLASSES = {}


def register_voxelnet(cls, name=None):
    name = cls.__name__ if name is None else name
    assert name not in REGISTERED_NETWORK_CLASSES, f"exist class: {REGISTERED_NETWORK_CLASSES}"
    REGISTERED_NETWORK_CLASSES[name] = cls
    return cls


def get_voxelnet_class(name):
    assert name in REGISTERED_NETWORK_CLASSES, f"available class: {REGISTERED_NETWORK_CLASSES}"
    return REGISTERED_NETWORK_CLASSES[name]


def create_cycle_constraint_data(xs, cat_dim=1):
    """all ordered pairs i<j (`voxel_odom_net.py:800-818`)."""
    assert len(xs) >= 2
    shape = xs[0].shape
    x1, x2 = [], []
    for i in range(len(xs)):
        for j in range(i + 1, len(xs)):
            x1.append(xs[i])
            x2.append(xs[j])
    return [torch.stack(x1, dim=cat_dim).reshape(-1, *shape[1:]), torch.stack(x2, dim=cat_dim).reshape(-1, *shape[1:])]


class _LazyOutputs(dict):
    """Output dict whose display-only entries are computed when somebody reads them (`ret["feature_mask"]`)."""
    lazy = {}

    def __missing__(self, key):
        fn = self.lazy.get(key)
        if fn is None:
            raise KeyError(key)
        val = self[key] = fn()
        return val

    def get(self, key, default=None):
        if key in self or key in self.lazy:
            return self[key]
        return default


def _detach_tree(x, to_cpu=False):
    if isinstance(x, torch.Tensor):
        x = x.detach()
        # device results are cloned: the head's outputs live in CUDA-graph static buffers that the next
        # call overwrites, while the reference hands out fresh tensors
        return x.cpu() if to_cpu else x.clone()
    if isinstance(x, (list, tuple)):
        return [_detach_tree(v, to_cpu) for v in x]
    if isinstance(x, dict):
        return {k: _detach_tree(v, to_cpu) for k, v in x.items()}
    return x


@register_voxelnet
class UnVoxelOdomNetICP3(nn.Module):
    def __init__(self, output_shape, pc_range=None, num_input_features=4, vfe_class_name="VoxelFeatureExtractor",
                 vfe_num_filters=(32, 128), with_distance=False, middle_class_name="SparseMiddleExtractor",
                 middle_num_input_features=-1, middle_num_filters_d1=(64,), middle_num_filters_d2=(64, 64),
                 middle_use_leakyReLU=False, middle_bn_type="BN", middle_relu_type="ReLU",
                 odom_class_name="ResNetOdomPred", odom_num_input_features=-1, odom_layer_nums=(3, 5, 5),
                 odom_layer_strides=(2, 2, 2), odom_num_filters=(128, 128, 256), odom_upsample_strides=(1, 2, 4),
                 odom_num_upsample_filters=(256, 256, 256), odom_pooling_type="avg_pool", odom_pooling_size=1,
                 odom_cycle_constraint=False, odom_conv_type="official", odom_format="rx+t",
                 odom_pred_pyramid_motion=False, odom_use_deep_supervision=False, odom_dense_predict=False,
                 odom_use_loss_mask=True, odom_use_dynamic_mask=False, odom_use_corr=False, odom_dropout=0.2,
                 odom_bn_type="BN", odom_conf_type="linear", odom_use_SPGN=False, odom_use_leakyReLU=False,
                 odom_first_conv_groups=1, odom_use_se=False, odom_use_sa=False, vfe_use_norm=True,
                 odom_enc_use_norm=True, odom_use_svd=False, odom_dropout_input=False, odom_cubic_pred_height=5,
                 freeze_bn=False, freeze_bn_affine=False, freeze_bn_start_step=1e20, sync_bn=False, use_GN=False,
                 encode_background_as_zeros=True, rotation_loss=None, translation_loss=None,
                 pyramid_rotation_loss=None, pyramid_translation_loss=None, consistency_loss=None,
                 measure_time=False, voxel_generator=None, pyloss_exp_w_base=0.5, testing=False, icp_iter=2,
                 name="voxel_odom_net", **kwargs):
        super().__init__()
        self.name = name
        self.testing = testing
        self._encode_background_as_zeros = encode_background_as_zeros
        self._num_input_features = num_input_features
        self.voxel_generator = voxel_generator
        self._rotation_loss = rotation_loss
        self._translation_loss = translation_loss
        self._pyramid_rotation_loss = pyramid_rotation_loss
        self._pyramid_translation_loss = pyramid_translation_loss
        self._consistency_loss = consistency_loss
        assert pyloss_exp_w_base > 0
        self._pyloss_exp_w_base = pyloss_exp_w_base
        assert icp_iter > 0, "The parameter of icp_iter should be larger than 0."
        self.icp_iter = icp_iter
        self.measure_time = measure_time
        self.voxel_feature_extractor = voxel_encoder.get_vfe_class(vfe_class_name)(
            num_input_features, vfe_use_norm, num_filters=vfe_num_filters, with_distance=with_distance,
            voxel_size=self.voxel_generator.voxel_size, pc_range=self.voxel_generator.point_cloud_range)
        self.middle_feature_extractor = middle.get_middle_class(middle_class_name)(
            output_shape, bn_type=middle_bn_type, use_GN=use_GN, sync_bn=sync_bn, use_leakyReLU=middle_use_leakyReLU,
            relu_type=middle_relu_type, num_input_features=middle_num_input_features,
            num_filters_down1=middle_num_filters_d1, num_filters_down2=middle_num_filters_d2)
        self.middle_feature_extractor_name = middle_class_name
        self.odom_predictor = odom_pred.get_odom_class(odom_class_name)(
            bn_type=odom_bn_type, enc_use_norm=odom_enc_use_norm, conv_type=odom_conv_type,
            layer_nums=odom_layer_nums, layer_strides=odom_layer_strides, num_filters=odom_num_filters,
            upsample_strides=odom_upsample_strides, num_upsample_filters=odom_num_upsample_filters,
            num_input_features=odom_num_input_features * 2, pooling_type=odom_pooling_type,
            pooling_size=odom_pooling_size, encode_background_as_zeros=True, use_groupnorm=use_GN, num_groups=32,
            dropout=odom_dropout, cycle_constraint=odom_cycle_constraint,
            pred_pyramid_motion=odom_pred_pyramid_motion, use_deep_supervision=odom_use_deep_supervision,
            use_loss_mask=odom_use_loss_mask, use_dynamic_mask=odom_use_dynamic_mask, odom_format=odom_format,
            point_cloud_range=pc_range, dense_predict=odom_dense_predict, use_correlation=odom_use_corr,
            conf_type=odom_conf_type, use_SPGN=odom_use_SPGN, use_leakyReLU=odom_use_leakyReLU,
            dropout_input=odom_dropout_input, first_conv_groups=odom_first_conv_groups, use_se=odom_use_se,
            use_sa=odom_use_sa, use_svd=odom_use_svd, cubic_pred_height=odom_cubic_pred_height,
            freeze_bn=freeze_bn, freeze_bn_affine=freeze_bn_affine, sync_bn=sync_bn, name="odomPred")
        self.freeze_bn = freeze_bn
        self.freeze_bn_affine = freeze_bn_affine
        self.freeze_bn_start_step = freeze_bn_start_step
        self.register_buffer("global_step", torch.LongTensor(1).zero_())
        self._step_host = None          # host mirror of global_step: no device read per query
        self.warm_flag = False
        self.__dict__["on_head_backward_done"] = None      # optional callback, see utils/distributed.FlatGradAllReducer
        self._time_dict, self._time_total_dict, self._time_count_dict = {}, {}, {}

    # ---- bookkeeping (voxel_odom_net.py:206-287) -------------------------------------------------
    def train(self, mode=True):
        super().train(mode)
        if self.freeze_bn and self.get_global_step() >= self.freeze_bn_start_step:
            for m in self.modules():
                if isinstance(m, nn.modules.batchnorm._BatchNorm):
                    m.eval()
                    if self.freeze_bn_affine:
                        if m.weight is not None:
                            m.weight.requires_grad = False
                        if m.bias is not None:
                            m.bias.requires_grad = False
        return self

    def start_timer(self, *names):
        if not self.measure_time:
            return
        torch.cuda.synchronize()
        for name in names:
            self._time_dict[name] = time.time()

    def end_timer(self, name):
        if not self.measure_time or name not in self._time_dict:
            return
        torch.cuda.synchronize()
        dt = time.time() - self._time_dict[name]
        self._time_count_dict[name] = self._time_count_dict.get(name, 0) + 1
        self._time_total_dict[name] = self._time_total_dict.get(name, 0.0) + dt
        self._time_dict[name] = 0

    def clear_timer(self):
        self._time_count_dict.clear()
        self._time_dict.clear()
        self._time_total_dict.clear()

    def get_avg_time_dict(self):
        return {name: val / max(1, self._time_count_dict[name]) for name, val in self._time_total_dict.items()}

    def update_global_step(self):
        self.global_step += 1
        if self._step_host is not None:
            self._step_host += 1

    def get_global_step(self):
        if self._step_host is None:
            self._step_host = int(self.global_step.cpu().numpy()[0])
        return self._step_host

    def clear_global_step(self):
        self.global_step.zero_()
        self._step_host = 0

    def _load_from_state_dict(self, *args, **kwargs):
        self._step_host = None
        return super()._load_from_state_dict(*args, **kwargs)

    def clear_metrics(self):
        pass

    # ---- tq target maps (voxel_odom_net.py:293-322) ----------------------------------------------
    def gen_tq_maps(self, odometries, spatial_size, pc_range, cubic_tq_map=False):
        if len(spatial_size) == 2:
            spatial_size = [1] + list(spatial_size)
        grid_size = np.array(list(spatial_size[::-1]))
        voxel_size = (pc_range[3:] - pc_range[0:3]) / grid_size
        spatial_size = grid_size if cubic_tq_map else grid_size[:2]
        origin_loc = ((0 - pc_range[0]) / (pc_range[3] - pc_range[0]) * grid_size[0],
                      (pc_range[4] - 0) / (pc_range[4] - pc_range[1]) * grid_size[1],
                      (0 - pc_range[2]) / (pc_range[5] - pc_range[2]) * grid_size[2])
        tq_maps = [generate_pointwise_local_transformation_tch(tq, spatial_size=spatial_size, origin_loc=origin_loc,
                                                               voxel_size=voxel_size, inv_trans_factor=-1)
                   for tq in odometries]
        return [torch.stack(tq_maps, dim=0)]

    # ---- forward ----------------------------------------------------------------------------------
    def _voxelize_on_device(self, points):
        vg = self.voxel_generator
        out = K.voxelize(points, vg.voxel_size, vg.point_cloud_range, vg.grid_size, max_points=vg.max_num_points,
                         max_voxels=vg.max_voxels_per_call, block_factor=vg.block_factor, block_size=vg.block_size,
                         height_threshold=vg.height_threshold, materialize=False, with_mean=True, with_table=True)
        # capacity-sized outputs; the live count stays on the device until the encoder's single count copy
        return out["mean"], out["coordinates"], out["num_points_per_voxel"], out["table"], out["n_dev"]

    # ---- ahead-of-time preparation (the reference voxelises in DataLoader workers, preprocess.py:493) ----
    def prepare(self, example, inputs_ready=True):
        """Voxelise `example["points"]` (device tensors, or pinned host tensors that are copied here) and
        build every index table of the sparse encoder on a side stream, including the one device->host
        copy of the row counts, so that the following `net(prepared)` neither waits for the main stream to
        drain nor leaves it idle.  `inputs_ready=False` makes the side stream wait for work already queued on
        the current stream (needed when the points were just produced there)."""
        assert "points" in example, "prepare() takes the raw-scan input form"
        dev = self.global_step.device
        st = self.__dict__.get("_prep_stream")
        if st is None:
            st = self.__dict__["_prep_stream"] = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        if not inputs_ready:
            st.wait_stream(main)
        with torch.cuda.stream(st):
            pts = [p.to(dev, non_blocking=True) if not p.is_cuda else p for p in example["points"]]
            voxels, coors, tables, n_devs = [], [], [], []
            for p in pts:
                m, c, _, tab, nd = self._voxelize_on_device(p)
                voxels.append(m)
                coors.append(c)
                tables.append(tab)
                n_devs.append(nd)
            finish = self.middle_feature_extractor.prepare_frames_begin(voxels, coors, tables, n_devs)
            ev = torch.cuda.Event()
            ev.record(st)
        # everything above was allocated on the preparation stream and will be read on the caller's stream
        made = list(pts) + voxels + coors + [t for t in n_devs if t is not None] + finish.pending.device_tensors()
        for tab in tables:
            if tab is not None:
                made += [tab.cells, tab.perm]
        seen = set()
        for t in made:
            if t is not None and t.is_cuda and id(t) not in seen:
                seen.add(id(t))
                t.record_stream(main)
        out = dict(example)
        # the row counts come back asynchronously: the tables are appended (and the step blocks on the counts, normally
        # long since there) when the prepared example is used, `finish()` in forward
        out["_prepared"] = {"finish": finish, "event": ev, "made": [t for t in made if t is not None and t.is_cuda],
                            "stream": main.cuda_stream}
        return out

    def network_forward(self, voxels, num_points, coors, batch_size, example):
        assert len(voxels) == len(num_points) == len(coors), "The lengths should be same."
        tables = example.get("_site_tables", [None] * len(voxels))
        n_devs = example.get("_n_dev", None)
        prepared = example.get("_prepared_frames", None)
        self.start_timer("voxel_feature_extractor")
        voxel_features = [self.voxel_feature_extractor(voxels[i], num_points[i], coors[i]) for i in range(len(voxels))]
        self.end_timer("voxel_feature_extractor")
        self.start_timer("middle forward")
        if hasattr(self.middle_feature_extractor, "forward_frames"):
            # all frames of the example share one pass through the sparse encoder
            spatial_features, middle_conf_preds, voxel_features, coors = self.middle_feature_extractor.forward_frames(
                voxel_features, coors, batch_size, tables, n_devs, prepared=prepared)
        else:
            spatial_features, middle_conf_preds = [], []
            for i in range(len(voxel_features)):
                ret, conf_pred = self.middle_feature_extractor(voxel_features[i], coors[i], batch_size)
                spatial_features.append(ret)
                middle_conf_preds.append(conf_pred)
        self.end_timer("middle forward")
        S = int(example.get("n_samples", 1))
        T = len(spatial_features) // S
        assert S * T == len(spatial_features), "n_samples must divide the number of frames"
        if S == 1:
            head_in = spatial_features
        else:       # frame slot t of every sample -> one [S,C,H,W] batch: all pairs of all samples in one head pass
            head_in = [torch.cat([spatial_features[s * T + t] for s in range(S)], dim=0) for t in range(T)]
        cb = self.__dict__.get("on_head_backward_done")
        if cb is not None and torch.is_grad_enabled() and head_in[0].requires_grad:
            # fires when autograd reaches the head's input, i.e. right after the head's backward has produced (and
            # accumulated) every head parameter gradient: lets a gradient reducer start on them early
            head_in[0].register_hook(lambda g, _cb=cb: (_cb(), g)[1])
        preds_dict = self.odom_predictor(head_in, tq_map_gt=None)
        if self.training or self.testing:
            # display maps of the training log (`voxel_odom_net.py:449-462`, consumed every display_step at
            # `train_hdf5.py:750`): ~40 small launches over the BEV maps.  With the reference's host_outputs they are
            # produced every step as there; with host_outputs=False they are produced on first access.
            def feature_mask(sf=spatial_features):
                with torch.no_grad():
                    return torch.cat(
                        [(torch.sum(torch.cat(sf[s * T:(s + 1) * T], dim=1), dim=1, keepdim=True) != 0).float()
                         for s in range(S)], dim=0)

            def middle_feature(sf=spatial_features):
                with torch.no_grad():
                    disp = [torch.mean(f.detach(), dim=1, keepdim=True) for f in sf]
                    return [(d - torch.min(d)) / (torch.max(d) - torch.min(d) + 1e-12) for d in disp]
            if example.get("host_outputs", True) or not self.training:
                preds_dict["feature_mask"] = feature_mask()
                preds_dict["middle_feature"] = middle_feature()
            else:
                preds_dict["_lazy_display"] = {"feature_mask": feature_mask, "middle_feature": middle_feature}
        preds_dict["middle_conf_preds"] = middle_conf_preds
        preds_dict["voxel_features"] = voxel_features
        preds_dict["voxel_coords"] = coors
        preds_dict["normal_preds"] = []
        return preds_dict

    def forward(self, example):
        if self.training:
            invalidate_weight_images()      # weights move every step, possibly through .data
        if ("voxels" in example and "points" not in example and "_prepared" not in example
                and example["voxels"][0].dim() == 3 and example["voxels"][0].shape[1] == 1
                and self.voxel_generator.max_num_points != 1):
            # pass-through layout of `_VoxelGenerator.generate(pass_through=True)`: voxels [P,1,F] ARE the raw scan in
            # scan order; voxelise on the device (bit-identical to voxelising up front, see tests)
            example = dict(example)
            example["points"] = [v[:, 0, :].contiguous() for v in example["voxels"]]
        if "_prepared" in example:
            prep = example["_prepared"]
            cur = torch.cuda.current_stream()
            cur.wait_event(prep["event"])
            if cur.cuda_stream != prep["stream"]:        # used on another stream than prepare() was told about
                for t in prep["made"]:
                    t.record_stream(cur)
            if "frames" not in prep:                     # first use: counts -> appended tables (on this stream)
                prep["frames"] = prep.pop("finish")()
            example = dict(example)
            example["_prepared_frames"] = prep["frames"]
            voxels, coors = prep["frames"]["features"], prep["frames"]["coors"]
            num_points = [None] * len(voxels)
            batch_size_dev = 1
        elif "points" in example:
            voxels, num_points, coors, tables, n_devs = [], [], [], [], []
            for pts in example["points"]:
                m, c, npts, tab, nd = self._voxelize_on_device(pts)
                voxels.append(m)
                coors.append(c)
                num_points.append(npts)
                tables.append(tab)
                n_devs.append(nd)
            example = dict(example)
            example["_site_tables"] = tables
            example["_n_dev"] = n_devs
            batch_size_dev = 1
        else:
            voxels, num_points, coors = example["voxels"], example["num_points"], example["coordinates"]
            if len(num_points[0].shape) == 2:       # padded multi-gpu layout (voxel_odom_net.py:480-507)
                vb, nb, cb = [], [], []
                for t in range(len(voxels)):
                    nv = example["num_voxels"][t].cpu().numpy().reshape(-1)
                    vb.append(torch.cat([voxels[t][i, :k] for i, k in enumerate(nv)], dim=0))
                    nb.append(torch.cat([num_points[t][i, :k] for i, k in enumerate(nv)], dim=0))
                    cb.append(torch.cat([coors[t][i, :k] for i, k in enumerate(nv)], dim=0))
                voxels, num_points, coors = vb, nb, cb
            batch_size_dev = example["num_voxels"][0].shape[0]
        preds_dict = self.network_forward(voxels, num_points, coors, batch_size_dev, example=example)
        if self.training:
            ret = self.loss(example, preds_dict)
            lazy = preds_dict.get("_lazy_display")
            if lazy is not None:
                ret = _LazyOutputs(ret)
                ret.lazy = lazy
            ret2 = {
                "t_conf": preds_dict["t_conf"], "r_conf": preds_dict["r_conf"],
                "pyramid_motion": preds_dict["pyramid_motion"], "dynamic_sigma": -1, "transformed_inputs": None,
                "tq_map_g": preds_dict["tq_map_g"], "local_motion": None, "down_masks": None,
                "middle_conf_preds": list(preds_dict["middle_conf_preds"]),
            }
            # the reference copies these ~10 maps to the host every step (voxel_odom_net.py:535-538);
            # `host_outputs=False` in the example keeps them on the device (detached)
            if lazy is None:
                ret2["middle_feature"], ret2["feature_mask"] = preds_dict["middle_feature"], preds_dict["feature_mask"]
            ret.update(_detach_tree(ret2, to_cpu=example.get("host_outputs", True)))
            return ret
        t_pred, r_pred = preds_dict["translation_preds"], preds_dict["rotation_preds"]
        if isinstance(t_pred, (list, tuple)):
            t_pred = t_pred[-1]
        if isinstance(r_pred, (list, tuple)):
            r_pred = r_pred[-1]
        out = {"translation_preds": _detach_tree(t_pred), "rotation_preds": _detach_tree(r_pred)}
        if self.testing:
            out["middle_conf_preds"] = [p.detach() for p in preds_dict["middle_conf_preds"]]
            out["voxel_features"] = [p.detach() for p in preds_dict["voxel_features"]]
            out["normal_preds"] = []
            out["tq_map_g"] = _detach_tree(preds_dict["tq_map_g"])
            out["pyramid_motion"] = _detach_tree(preds_dict["pyramid_motion"])
            out["t_conf"] = _detach_tree(preds_dict["t_conf"])
            out["r_conf"] = _detach_tree(preds_dict["r_conf"])
            out["normal_gt"] = example.get("normal_gt", None)
        return out

    # ---- loss (voxel_odom_net.py:324-376, 586-798) -------------------------------------------------
    def loss(self, example, preds_dict):
        T_preds, R_preds = preds_dict["translation_preds"], preds_dict["rotation_preds"]
        dtype = T_preds[0].dtype
        self.start_timer("Create_loss forward")
        pyramid_loss = torch.zeros([1], dtype=dtype, device=T_preds[0].device)
        translation_loss, rotation_loss, pyramid_T_losses, pyramid_R_losses, C_loss = self.create_loss(
            preds_dict, example, self._translation_loss, self._rotation_loss,
            pyramid_rotation_loss=self._pyramid_rotation_loss,
            pyramid_translation_loss=self._pyramid_translation_loss, consistency_loss=self._consistency_loss)
        pyramid_num = len(pyramid_T_losses)
        l8 = self.__dict__.pop("_losses8", None)
        if l8 is not None and pyramid_num == 3:
            # fused loss tail: loss = T + R + sum_i base^(n-i) (pyT_i + pyR_i) + C as one weighted sum of its [8] output
            w8 = self.__dict__.get("_loss_w8")
            if w8 is None or w8.device != l8.device:
                pw = [self._pyloss_exp_w_base ** (pyramid_num - i) for i in range(pyramid_num)]
                w8 = self.__dict__["_loss_w8"] = torch.tensor([1.0, 1.0] + pw + pw, dtype=l8.dtype, device=l8.device)
            loss = (l8 * w8).sum().reshape(1) + C_loss
            with torch.no_grad():
                pyramid_loss = (l8[2:] * w8[2:]).sum().reshape(1)
        else:
            for i, (t_loss, r_loss) in enumerate(zip(pyramid_T_losses, pyramid_R_losses)):
                pyramid_loss = pyramid_loss + self._pyloss_exp_w_base ** (pyramid_num - i) * (t_loss + r_loss)
            loss = translation_loss + rotation_loss + pyramid_loss + C_loss
        self.end_timer("Create_loss forward")
        return {"loss": loss, "translation_loss": translation_loss.detach(), "rotation_loss": rotation_loss.detach(),
                "pyramid_loss": pyramid_loss.detach(), "C_loss": C_loss.detach(),
                "translation_preds": _detach_tree(T_preds[0]), "rotation_preds": _detach_tree(R_preds[0])}

    def create_loss(self, preds_dict, example, translation_loss, rotation_loss, pyramid_translation_loss=None,
                    pyramid_rotation_loss=None, pyramid_preds=None, consistency_loss=None):
        translation_preds, rotation_preds = preds_dict["translation_preds"], preds_dict["rotation_preds"]
        if not isinstance(translation_preds, (list, tuple)):
            translation_preds = [translation_preds]
        if not isinstance(rotation_preds, (list, tuple)):
            rotation_preds = [rotation_preds]
        dtype, device = translation_preds[0].dtype, translation_preds[0].device
        pyramid_preds = preds_dict["pyramid_motion"]
        step = self.get_global_step()
        if "icp_odometry" in example:
            icp = example["icp_odometry"].view(-1, 7)
        else:                                   # the shipped dataset fills it with zeros (preprocess.py:618)
            icp = torch.zeros((translation_preds[0].shape[0], 7), dtype=dtype, device=device)
        translation_targets, rotation_targets = icp[:, :3], icp[:, 3:]
        if translation_loss._loss_weight == 0:
            self.warm_flag = True
        if self.warm_flag:
            warm_weight = 1.0 / (0.001 * step + 1) if step < 1500 else 0
            translation_loss._loss_weight = warm_weight
            rotation_loss._loss_weight = warm_weight
        else:
            warm_weight = 0

        C_loss = torch.zeros([1], dtype=dtype, device=device)
        res_r, res_t = None, None
        if consistency_loss is not None:
            assert len(preds_dict["middle_conf_preds"]) > 0
            feats = preds_dict["voxel_features"]
            S = int(example.get("n_samples", 1))
            T = len(feats) // S
            n_pairs = T * (T - 1) // 2
            cols = [0, 1, 2, 4, 5, 6] if feats[0].shape[1] > 6 else [0, 1, 2, 3, 4, 5]
            weights = [0.01, 0.01, 0.05, 0.1, 1]
            res_rs, res_ts = [], []
            # the samples of a step are independent: each keeps its own common point count (voxel_odom_net.py:646-651)
            # two-frame samples (the shipped training / evaluation setup): the pair's tensors are views of the frames'
            # rows and the predicted pose is applied by one kernel (csrc/pair_transform.cu) instead of ~80 torch ops
            fast = (T == 2 and len(rotation_preds) == 1 and rotation_preds[0].shape[-1] == 4 and feats[0].shape[1] > 6
                    and feats[0].is_cuda)
            for smp in range(S if fast else 0):
                f0, f1 = feats[smp * T], feats[smp * T + 1]
                c0, c1 = preds_dict["middle_conf_preds"][smp * T], preds_dict["middle_conf_preds"][smp * T + 1]
                n = min(f0.shape[0], f1.shape[0])
                p0, p1 = f0[:n], f1[:n]
                target, R_pred = K.pair_transform(p1, rotation_preds[0][smp], translation_preds[0][smp],
                                                  identity=step <= 1500)
                icp_iter = self.icp_iter if step > 1500 else 5
                l, res_r, res_t = consistency_loss(
                    p0[None, :, :3], target[None], cov_pred=c0[None, :n], cov_target=c1[None, :n], R_pred=R_pred[None],
                    t_pred=None, normal_pred=p0[None, :, 4:7], normal_target=None, mask=None, icp_iter=icp_iter)
                C_loss = C_loss + l * ((1 - warm_weight) * weights[-1] / S)
                res_rs.append(res_r)
                res_ts.append(res_t)
            for smp in range(0 if fast else S):
                fr = slice(smp * T, (smp + 1) * T)
                pr = slice(smp * n_pairs, (smp + 1) * n_pairs)
                points = [[p[:, cols][None]] for p in feats[fr]]
                point_confs = [p[None] for p in preds_dict["middle_conf_preds"][fr]]
                min_len = min(p[0].shape[1] for p in points)
                points = [[p[0][:, :min_len]] for p in points]
                point_confs = create_cycle_constraint_data([p[:, :min_len] for p in point_confs])
                new_points = []
                for h, _ in enumerate(points[0]):
                    new_points.append(create_cycle_constraint_data([points[t][h] for t in range(len(points))], 1))
                if len(new_points) < len(rotation_preds):
                    new_points = new_points + [new_points[-1]] * (len(rotation_preds) - len(new_points))
                else:
                    new_points = new_points[:len(rotation_preds)]
                for i, (R_pred, T_pred, weight) in enumerate(zip(rotation_preds, translation_preds,
                                                                 weights[-len(translation_preds):])):
                    R_pred, T_pred = R_pred[pr], T_pred[pr]
                    if R_pred.shape[-1] == 9:
                        R_pred = R_pred.reshape(-1, 3, 3)
                    else:
                        R_pred = pose_utils.quaternion_to_rotation_matrix(roll(R_pred, shift=-1, dim=-1))
                    if step <= 1500:
                        R_pred = torch.eye(3, device=device, dtype=dtype).expand(R_pred.shape[0], 3, 3).contiguous()
                        T_pred = torch.zeros_like(T_pred)
                    p0, p1 = new_points[-(i + 1)][0], new_points[-(i + 1)][1]
                    transformed_p1_gt = p1[:, :, :3] @ R_pred.transpose(1, 2) + T_pred[:, None, :]
                    transformed_p1 = p0[:, :, :3]
                    transformed_normal1_gt = p1[:, :, 3:] @ R_pred.detach().transpose(1, 2)
                    transformed_normal1 = p0[:, :, 3:]
                    icp_iter = self.icp_iter if step > 1500 else 5
                    l, res_r, res_t = consistency_loss(
                        transformed_p1, transformed_p1_gt, cov_pred=point_confs[0], cov_target=point_confs[1],
                        R_pred=R_pred, t_pred=T_pred, normal_pred=transformed_normal1.detach(),
                        normal_target=transformed_normal1_gt.detach(), mask=None, icp_iter=icp_iter)
                    C_loss = C_loss + l * ((1 - warm_weight) * weight / S)
                res_rs.append(res_r)
                res_ts.append(res_t)
            res_r, res_t = (res_rs[0], res_ts[0]) if S == 1 else (torch.cat(res_rs), torch.cat(res_ts))

        if (res_r is not None and len(pyramid_preds) > 0 and len(translation_preds) == 1
                and pyramid_translation_loss is not None and pyramid_rotation_loss is not None):
            # everything downstream of the ICP result has shapes fixed by the BEV grid: pseudo labels,
            # target (t,q) maps, pose and pyramid losses run as one captured CUDA graph (fwd + bwd)
            outs = self._loss_tail(translation_preds[0], rotation_preds[0], pyramid_preds, res_r, res_t,
                                   identity_pose=step <= 1500)
            n_py = len(pyramid_preds)
            T_loss, R_loss = outs[0], outs[1]
            pyramid_T_losses, pyramid_R_losses = list(outs[2:2 + n_py]), list(outs[2 + n_py:2 + 2 * n_py])
            example["tq_maps"] = [outs[2 + 2 * n_py]]
            return T_loss, R_loss, pyramid_T_losses, pyramid_R_losses, C_loss

        if res_r is not None and res_t is not None:
            R_all, T_all = rotation_preds[-1], translation_preds[-1]
            R_all = R_all.reshape(-1, 3, 3) if R_all.shape[-1] == 9 else \
                pose_utils.quaternion_to_rotation_matrix(roll(R_all, shift=-1, dim=-1))
            if step <= 1500:
                R_all = torch.eye(3, device=device, dtype=dtype).expand(R_all.shape[0], 3, 3).contiguous()
                T_all = torch.zeros_like(T_all)
            rotation_targets, translation_targets = self._pseudo_labels(res_r, res_t, R_all, T_all)

        if len(pyramid_preds) > 0:
            tq_map_targets = self.gen_tq_maps(
                torch.cat([translation_targets, rotation_targets], dim=-1).reshape(-1, 7),
                spatial_size=pyramid_preds[-1][0].shape[2:], pc_range=self.odom_predictor.point_cloud_range,
                cubic_tq_map=self.odom_predictor._cubic_pred_height > 0)
            example["tq_maps"] = tq_map_targets
        pyramid_targets = list(example["tq_maps"])

        T_loss = 0
        for p in translation_preds:
            T_loss = T_loss + translation_loss(p, translation_targets)
        R_loss = 0
        for p in rotation_preds:
            R_loss = R_loss + rotation_loss(p, rotation_targets)
        if pyramid_translation_loss is None or pyramid_rotation_loss is None:
            return T_loss, R_loss
        pyramid_T_losses, pyramid_R_losses = self._pyramid_losses(pyramid_preds, pyramid_targets[0],
                                                                  pyramid_translation_loss, pyramid_rotation_loss)
        return T_loss, R_loss, pyramid_T_losses, pyramid_R_losses, C_loss

    # ---- pieces of create_loss downstream of the ICP result (voxel_odom_net.py:727-795) ------------
    @staticmethod
    def _pseudo_labels(res_r, res_t, R_pred, T_pred):
        rotation_targets = res_r @ R_pred.detach()
        rotation_targets = pose_utils.rotation_matrix_to_quaternion(rotation_targets)
        rotation_targets = roll(rotation_targets, 1, dim=-1)
        rotation_targets = rotation_targets * torch.sign(rotation_targets[:, 0:1])
        translation_targets = (res_r @ T_pred[..., None].detach() + res_t[..., None]).squeeze(-1)
        return rotation_targets, translation_targets

    @staticmethod
    def _pyramid_losses(pyramid_preds, target_map, pyramid_translation_loss, pyramid_rotation_loss):
        pyramid_T_losses, pyramid_R_losses = [], []
        for i, _ in enumerate(pyramid_preds):
            T_pred_i, R_pred_i = pyramid_preds[i][0][:, :3], pyramid_preds[i][0][:, 3:]
            pred_mask = pyramid_preds[i][1]
            T_target, R_target = target_map[:, :3], target_map[:, 3:]
            if T_target.shape != T_pred_i.shape:
                T_target = F.interpolate(T_target, size=T_pred_i[0, 0].shape, mode="nearest")
            if R_target.shape != R_pred_i.shape:
                R_target = F.interpolate(R_target, size=T_pred_i[0, 0].shape, mode="nearest")
            pyramid_T_losses.append(pyramid_translation_loss(T_pred_i, T_target, mask=pred_mask[:, :1]))
            pyramid_R_losses.append(pyramid_rotation_loss(R_pred_i, R_target, mask=pred_mask[:, -1:]))
        return pyramid_T_losses, pyramid_R_losses

    def _loss_tail_eager(self, T_pred, q_pred, pyramid_flat, res_r, res_t, identity_pose):
        """(T_pred [B,3], q_pred [B,4], [pred_0, mask_0, pred_1, ...], res_r, res_t) ->
        (T_loss, R_loss, pyT_0.., pyR_0.., tq_map_target)."""
        pyramid_preds = [[pyramid_flat[2 * i], pyramid_flat[2 * i + 1]] for i in range(len(pyramid_flat) // 2)]
        R_pred = pose_utils.quaternion_to_rotation_matrix(roll(q_pred, shift=-1, dim=-1))
        T_used = T_pred
        if identity_pose:                                   # step <= 1500 (voxel_odom_net.py:677-679)
            R_pred = torch.eye(3, device=T_pred.device, dtype=T_pred.dtype).expand(R_pred.shape[0], 3, 3)
            T_used = torch.zeros_like(T_pred)
        rotation_targets, translation_targets = self._pseudo_labels(res_r, res_t, R_pred, T_used)
        tq_map = self.gen_tq_maps(torch.cat([translation_targets, rotation_targets], dim=-1).reshape(-1, 7),
                                  spatial_size=pyramid_preds[-1][0].shape[2:],
                                  pc_range=self.odom_predictor.point_cloud_range,
                                  cubic_tq_map=self.odom_predictor._cubic_pred_height > 0)[0]
        T_loss = self._translation_loss(T_pred, translation_targets)
        R_loss = self._rotation_loss(q_pred, rotation_targets)
        pyT, pyR = self._pyramid_losses(pyramid_preds, tq_map, self._pyramid_translation_loss,
                                        self._pyramid_rotation_loss)
        return (T_loss, R_loss, *pyT, *pyR, tq_map)

    def _loss_tail(self, T_pred, q_pred, pyramid_preds, res_r, res_t, identity_pose):
        """pseudo labels -> target maps -> pose + pyramid losses: one fused kernel forward, one backward
        (layers/pose_tail.py over csrc/pose_tail.cu) when the configuration is the shipped one; otherwise the same
        steps as torch ops."""
        flat = []
        for pred, mask in pyramid_preds:
            flat += [pred, mask]
        mods = (self._translation_loss, self._rotation_loss, self._pyramid_translation_loss, self._pyramid_rotation_loss)
        fused = (odom_pred.USE_OWN_TAIL and T_pred.is_cuda and len(pyramid_preds) == 3
                 and all(getattr(m, "focal_gamma", 1) == 0 for m in mods)
                 and not (self.odom_predictor._cubic_pred_height > 0)
                 and pyramid_preds[0][0].shape[2] * 4 == pyramid_preds[2][0].shape[2]
                 and pyramid_preds[1][0].shape[2] * 2 == pyramid_preds[2][0].shape[2])
        if not fused:
            return self._loss_tail_eager(T_pred, q_pred, flat, res_r, res_t, identity_pose)
        H, W = pyramid_preds[2][0].shape[2:]
        geom = self.__dict__.get("_loss_geom")
        if geom is None or (geom.H, geom.W) != (H, W):
            geom = self.__dict__["_loss_geom"] = loss_geometry(H, W, self.odom_predictor.point_cloud_range)
        l8, tq_map = loss_tail(T_pred, q_pred, pyramid_preds, res_r, res_t, [m.alpha for m in mods],
                               [m._loss_weight for m in mods], identity_pose, geom)
        self.__dict__["_losses8"] = l8          # picked up by loss(): one weighted sum instead of eight scalar chains
        return (*[l8[i:i + 1] for i in range(8)], tq_map)
