"""`rslo/builder/losses_builder.py:23-151`."""
from ..core import losses


def _build_pose_loss(cfg):
    if cfg.loss_type == "AdaptiveWeightedL2":
        if cfg.balance_scale <= 0:
            cfg.balance_scale = 1
        return losses.AdaptiveWeightedL2Loss(cfg.init_alpha, learn_alpha=not cfg.not_learn_alpha,
                                             loss_weight=cfg.weight, focal_gamma=cfg.focal_gamma,
                                             balance_scale=cfg.balance_scale)
    raise ValueError(f"loss_type {cfg.loss_type!r} is outside the shipped configs")


def _build_consistency_loss(cfg):
    if cfg.loss_type == "Aleat5_1ChamferL2NormalWeightedALLSVDLoss":
        assert cfg.penalize_ratio > 0 and cfg.pred_downsample_ratio > 0 and cfg.reg_weight > 0 and cfg.sph_weight > 0
        return losses.Aleat5_1ChamferL2NormalWeightedALLSVDLoss(
            loss_weight=cfg.weight, penalize_ratio=cfg.penalize_ratio, sample_block_size=cfg.sample_block_size,
            norm=cfg.norm, pred_downsample_ratio=cfg.pred_downsample_ratio, reg_weight=cfg.reg_weight,
            sph_weight=cfg.sph_weight)
    if cfg.loss_type == "":
        return None
    raise ValueError(f"consistency loss_type {cfg.loss_type!r} is outside the shipped configs")


def build(loss_config):
    """-> (rotation, translation, pyramid_rotation, pyramid_translation, consistency); when the pyramid
    losses are not configured they ARE the main loss modules (shared alpha), `losses_builder.py:40-50`."""
    rot = _build_pose_loss(loss_config.rotation_loss)
    trans = _build_pose_loss(loss_config.translation_loss)
    py_rot = _build_pose_loss(loss_config.pyramid_rotation_loss) if loss_config.pyramid_rotation_loss.loss_type != "" else rot
    py_trans = _build_pose_loss(loss_config.pyramid_translation_loss) if loss_config.pyramid_translation_loss.loss_type != "" else trans
    cons = _build_consistency_loss(loss_config.consistency_loss)
    assert loss_config.rigid_transform_loss.weight == 0, "rigid_transform_loss is outside the shipped configs"
    return rot, trans, py_rot, py_trans, cons
