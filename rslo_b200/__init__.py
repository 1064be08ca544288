"""rslo_b200 — B200-native implementation of RSLO's per-frame-pair hot path.

Host code mirrors the reference's `rslo.models` / `rslo.layers` / builder surface; compute runs in
hand-written sm_100a CUDA kernels behind the C ABI in include/rslo_b200.h (rslo_b200/_C).
"""
__version__ = "0.1.0"

import os as _os

# No process-wide backend switches here: the path's own kernels are FP32-exact by construction (split-TF32),
# torch's matmul TF32 is off by default, and the cuDNN A/B trunk scopes its flags itself (models/odom_pred.py).

DEFAULT_CONFIG = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "config", "kitti_ours.prototxt")


def build_network(config_path=None, testing=False, measure_time=False, seed=None):
    """prototxt -> (net, voxel_generator), the two builder calls `train_hdf5.py:92-101` makes."""
    import torch

    from .builder import config as _config
    from .builder import second_builder, voxel_builder
    cfg = _config.load(config_path or DEFAULT_CONFIG)
    vg = voxel_builder.build(cfg.model.second.voxel_generator)
    # preprocess.max_number_of_voxels of the input reader that feeds this net (dataset_builder.py:78, preprocess.py:493):
    # the eval reader's when testing (evaluate.py builds its dataset from eval_input_reader), else the train reader's
    for reader in (("eval_input_reader", "train_input_reader") if testing else ("train_input_reader", "eval_input_reader")):
        try:
            vg.max_voxels_per_call = int(getattr(cfg, reader).preprocess.max_number_of_voxels)
            break
        except (AttributeError, TypeError):
            continue
    if seed is not None:
        torch.manual_seed(seed)
    net = second_builder.build(cfg.model.second, vg, measure_time=measure_time, testing=testing)
    return net, vg
