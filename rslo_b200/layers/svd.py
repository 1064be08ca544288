"""`rslo/layers/svd.py:13-64`: SVDHead()(src [B,3,n], tgt [B,3,n], weight [B,n]) -> (R [B,3,3], t [B,3])."""
import torch
from torch import nn

from .. import kernels as K


class SVDHead(nn.Module):
    """Weighted Kabsch through csrc/kabsch.cu (double-precision moments + on-device 3x3 Jacobi SVD):
    no cuSOLVER call and no `det < 0` host branch.  Inputs are treated as constants (the reference
    feeds it detached tensors, `losses.py:441-447`)."""

    def __init__(self, args=None):
        super().__init__()
        self.reflect = nn.Parameter(torch.eye(3), requires_grad=False)
        self.reflect[2, 2] = -1

    @torch.no_grad()
    def forward(self, src, tgt, weight=None):
        Rs, ts = [], []
        for b in range(src.size(0)):
            R, t = K.kabsch(src[b].t().contiguous(), tgt[b].t().contiguous(),
                            None if weight is None else weight[b].contiguous())
            Rs.append(R)
            ts.append(t)
        return torch.stack(Rs), torch.stack(ts)
