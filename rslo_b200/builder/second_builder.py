"""`rslo/builder/second_builder.py:26-136`: model config -> network."""
from ..models.voxel_odom_net import get_voxelnet_class
from . import losses_builder


def build(model_cfg, voxel_generator, measure_time=False, testing=False):
    vfe_num_filters = list(model_cfg.voxel_feature_extractor.num_filters)
    grid_size = voxel_generator.grid_size
    dense_shape = [1] + grid_size[::-1].tolist() + [vfe_num_filters[-1]]
    rot, trans, py_rot, py_trans, cons = losses_builder.build(model_cfg.loss)
    m, o = model_cfg.middle_feature_extractor, model_cfg.odom_predictor
    return get_voxelnet_class(model_cfg.network_class_name)(
        dense_shape, pc_range=voxel_generator.point_cloud_range,
        vfe_class_name=model_cfg.voxel_feature_extractor.module_class_name, vfe_num_filters=vfe_num_filters,
        middle_class_name=m.module_class_name, middle_num_input_features=m.num_input_features,
        middle_num_filters_d1=list(m.num_filters_down1), middle_num_filters_d2=list(m.num_filters_down2),
        middle_use_leakyReLU=m.use_leakyReLU, middle_relu_type=m.relu_type,
        odom_class_name=o.module_class_name, odom_num_input_features=o.num_input_features,
        odom_layer_nums=o.layer_nums, odom_layer_strides=o.layer_strides, odom_num_filters=o.num_filters,
        odom_upsample_strides=o.upsample_strides, odom_num_upsample_filters=o.num_upsample_filters,
        odom_pooling_size=o.pool_size, odom_pooling_type=o.pool_type, odom_cycle_constraint=o.cycle_constraint,
        odom_conv_type=o.conv_type, odom_format=o.odom_format, odom_pred_pyramid_motion=o.pred_pyramid_motion,
        odom_use_deep_supervision=o.use_deep_supervision, odom_use_loss_mask=not o.not_use_loss_mask,
        odom_use_dynamic_mask=o.use_dynamic_mask, odom_dense_predict=o.dense_predict, odom_use_corr=o.use_corr,
        odom_dropout=o.dropout, odom_conf_type=o.conf_type, odom_use_SPGN=o.use_SPGN,
        odom_use_leakyReLU=o.use_leakyReLU, vfe_use_norm=not model_cfg.voxel_feature_extractor.not_use_norm,
        middle_bn_type=m.bn_type, odom_bn_type=o.bn_type, odom_enc_use_norm=not o.not_use_enc_norm,
        odom_dropout_input=o.dropout_input, odom_first_conv_groups=max(1, o.first_conv_groups),
        odom_use_se=o.odom_use_se, odom_use_sa=o.odom_use_sa, odom_use_svd=o.use_svd,
        odom_cubic_pred_height=o.cubic_pred_height, freeze_bn=model_cfg.freeze_bn,
        freeze_bn_affine=model_cfg.freeze_bn_affine, freeze_bn_start_step=model_cfg.freeze_bn_start_step,
        use_GN=model_cfg.use_GN, num_input_features=model_cfg.num_point_features,
        encode_background_as_zeros=model_cfg.encode_background_as_zeros,
        with_distance=model_cfg.voxel_feature_extractor.with_distance, rotation_loss=rot, translation_loss=trans,
        pyramid_rotation_loss=py_rot, pyramid_translation_loss=py_trans, consistency_loss=cons,
        measure_time=measure_time, voxel_generator=voxel_generator, pyloss_exp_w_base=model_cfg.loss.pyloss_exp_w_base,
        testing=testing, icp_iter=model_cfg.icp_iter)
