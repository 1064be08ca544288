"""Optimizer side of the training step (SURVEY §8 f-N2): fused clip + weight decay + Adam over the flat gradient
buffer, and the one-cycle schedule that drives its lr / momentum."""
from .fused_optim import FusedAdamClip, OneCycle, annealing_cos, build_optimizer  # noqa: F401
