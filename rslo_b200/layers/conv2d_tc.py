"""`nn.Conv2d` of the dense head (`rslo/models/odom_pred_base.py:155-276`, `rslo/layers/MaskConv.py:20-73`).

The module only holds the parameters under the reference's names and OIHW shapes.  On the GPU the head's trunk
never calls it: `rslo_b200/layers/head_tc.py` walks the module tree and runs every convolution on
csrc/conv2d_tc.cu (TMA-staged split-TF32 tcgen05 implicit GEMM; forward, data gradient, weight gradient).
`forward` (plain torch) remains for the A/B switch RSLO_HEAD_TC=0 and for the float64 evaluation in the tests.
"""
from torch import nn


class Conv2dTC(nn.Conv2d):
    pass
