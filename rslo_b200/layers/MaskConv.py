"""Mask-propagating 2-D convolution (`rslo/layers/MaskConv.py:20-73`)."""
import torch
from torch import nn

from .conv2d_tc import Conv2dTC


class MaskConv(nn.Module):
    """conv(bias=False) on the features in parallel with a max-pool of the occupancy mask; returns
    ``[features, mask]``.  The conv is held as ``.conv1`` so state_dict keys read
    ``...conv1.conv1.weight`` as in the reference.

    ``propagate_mask=False`` skips the mask pooling and returns ``[features, None]``: in the shipped
    head the propagated masks are collected and never read ("dead weight"), so
    the model builds its MaskConvs that way; the operator default keeps the reference behaviour."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True, max_pool_mask=True,
                 groups=1, propagate_mask=True):
        super().__init__()
        assert max_pool_mask, "only the max-pool mask variant is used by the shipped configs"
        self.out_channels = out_channels
        self.use_bias = bias
        self.conv1 = Conv2dTC(in_channels, out_channels, kernel_size=kernel_size, stride=stride, bias=False,
                               padding=padding, groups=groups)
        self.max_pool_mask = max_pool_mask
        self.propagate_mask = propagate_mask
        self.mask_pool = nn.MaxPool2d(kernel_size, stride=stride, padding=padding)
        self.normalize_const = 1

    def forward(self, x):
        if not isinstance(x, (list, tuple)):
            mask = (torch.sum(x.abs(), dim=1, keepdim=True) != 0).float().detach() if self.propagate_mask else None
            x = [x, mask]
        tensor, mask = x
        tensor = self.conv1(tensor)
        if self.propagate_mask and mask is not None:
            mask = self.mask_pool(mask).detach()
        else:
            mask = None
        return [tensor, mask]
