"""Tuple-aware normalisation / activation layers (`rslo/layers/SparseConv.py:96-132` and the
SPC_ReLU / SPC_LeakyReLU / SPC_BN2d classes further down that file): they accept either a tensor or
``[tensor, mask]`` and pass the mask through."""
from torch import nn


class SPC_BN2d(nn.BatchNorm2d):
    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True,
                 process_group=None, channel_last=False, fuse_relu=False, noise_scale_std=0, noise_shift_std=0):
        super().__init__(num_features, eps, momentum, affine, track_running_stats)
        assert noise_scale_std == 0 and noise_shift_std == 0, "BN noise is not used by the shipped configs"

    def forward(self, x):
        if isinstance(x, (tuple, list)):
            return [super().forward(x[0]), x[1]]
        return super().forward(x)


class SPC_SyncBN2d(SPC_BN2d):
    """The reference subclasses apex.parallel.SyncBatchNorm (eps 1e-3, momentum 0.01 set by the
    head, `odom_pred_base.py:140-141`).  Here it is a torch BatchNorm2d subclass (so the optimizer's
    BN/non-BN parameter split by isinstance still works, `fastai_optim.py:11-25`) with per-rank
    statistics: north_star all-reduces gradients only.  Identical at world size 1 and in eval."""


class SPC_ReLU(nn.ReLU):
    def forward(self, x):
        if isinstance(x, (tuple, list)):
            return [super().forward(x[0]), x[1]]
        return super().forward(x)


class SPC_LeakyReLU(nn.LeakyReLU):
    def forward(self, x):
        if isinstance(x, (tuple, list)):
            return [super().forward(x[0]), x[1]]
        return super().forward(x)
