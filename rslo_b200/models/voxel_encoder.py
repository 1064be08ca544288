"""Voxel feature extractors (registry surface of `rslo/models/voxel_encoder.py:11-26`)."""
from torch import nn

from .. import kernels as K

REGISTERED_VFE_CLASSES = {}


def register_vfe(cls, name=None):
    name = cls.__name__ if name is None else name
    assert name not in REGISTERED_VFE_CLASSES, f"exist class: {REGISTERED_VFE_CLASSES}"
    REGISTERED_VFE_CLASSES[name] = cls
    return cls


def get_vfe_class(name):
    assert name in REGISTERED_VFE_CLASSES, f"available class: {REGISTERED_VFE_CLASSES}"
    return REGISTERED_VFE_CLASSES[name]


@register_vfe
class SimpleVoxel_XYZINormalC(nn.Module):
    """Per-voxel mean of the stored points with the normal (cols 4:7) renormalised
    (`rslo/models/voxel_encoder.py:258-280`).  Parameter free; one kernel (csrc/voxelize.cu
    k_vfe_mean).  When the voxeliser already produced the means (fused path) they are passed
    through as `features` of shape [N,7]."""

    def __init__(self, num_input_features=8, use_norm=True, num_filters=(32, 128), with_distance=False,
                 voxel_size=(0.2, 0.2, 4), pc_range=(0, -40, -3, 70.4, 40, 1), name="VoxelFeatureExtractor"):
        super().__init__()
        self.name = name
        self.num_input_features = num_input_features

    def forward(self, features, num_voxels, coors):
        if features.dim() == 2:          # fused voxeliser+VFE output
            return features
        assert self.num_input_features == 7, "kernel computes the 7-feature XYZI+normal mean"
        return K.vfe_mean(features, num_voxels.int())
