"""Confidence head (`rslo/layers/confidence.py:5-38`): conv stack -> masked spatial softmax."""
import torch
from torch import nn
import torch.nn.functional as F


class ConfidenceModule(nn.Module):
    def __init__(self, conf_model, conf_type="softmax"):
        super().__init__()
        assert conf_type in ["linear", "softmax"]
        self.conf_model = conf_model
        self.conf_type = conf_type
        self.softmax = nn.Softmax(dim=-1)

    def forward(self, x, extra_mask=None, temperature=1, return_logit=False, logit=None):
        """``logit`` lets the caller reuse the conv stack's output when the same input is scored at
        a second temperature (the reference recomputes it, `odom_pred.py:242-258`)."""
        if logit is None:
            logit = self.conf_model(x)
        if extra_mask is None:
            extra_mask = torch.ones_like(logit)
        if self.conf_type == "linear":
            conf = (F.elu(logit) + 1 + 1e-12) * (extra_mask + 1e-12)
        else:
            conf = torch.where(extra_mask > 0, logit, torch.full_like(logit, -1000))
            shape = conf.shape
            conf = self.softmax(conf.reshape(*shape[0:2], -1) / temperature).reshape(*shape)
        if return_logit:
            return conf, logit
        return conf
