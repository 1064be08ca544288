"""`rslo/layers/common.py`: ParameterLayer (the head's `dynamic_sigma`)."""
import torch
from torch import nn


class ParameterLayer(nn.Module):
    def __init__(self, init_value, requires_grad=True):
        super().__init__()
        self.param = nn.Parameter(torch.as_tensor(init_value).clone().float(), requires_grad=requires_grad)

    def forward(self, *a, **k):
        return self.param
