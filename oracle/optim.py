"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference's optimizer step for SURVEY §8 row f-N2.

Restates (torch tensor ops on the CPU, one parameter at a time):
  * torch.nn.utils.clip_grad_norm_(parameters, max_norm)             as called at /root/reference/train_hdf5.py:671
  * OptimWrapper.step() with true_wd / bn_wd                         /root/reference/rslo/torchplus/train/fastai_optim.py:181-194
  * torch.optim.Adam(betas=(mom, 0.99), eps=1e-8, weight_decay=0)    built at rslo/builder/optimizer_builder.py:101-118
  * OneCycle                                                         rslo/torchplus/train/learning_schedules_fastai.py:44-95
Pinned by tests/test_cpu_oracle.py::test_optimizer_oracle_matches_reference_live against the reference's own
OptimWrapper + OneCycle classes (imported through oracle/ref_shim.py) around the installed torch.optim.Adam.
"""
import math

import torch


def clip_grad_norm(grads, max_norm):
    """-> (total_norm, list of clipped grads); `None` entries (no gradient) are skipped."""
    present = [g for g in grads if g is not None]
    total = torch.norm(torch.stack([torch.norm(g.detach(), 2.0) for g in present]), 2.0)
    coef = max_norm / (total + 1e-6)
    if float(coef) < 1:
        return total, [None if g is None else g * coef for g in grads]
    return total, list(grads)


def adam_step(params, grads, state, lr, beta1, beta2=0.99, eps=1e-8, wd=0.0, true_wd=True):
    """One OptimWrapper.step(): state = {"step": [..], "m": [..], "v": [..]} per parameter; params updated in place."""
    for i, (p, g) in enumerate(zip(params, grads)):
        if true_wd:
            p.mul_(1 - wd * lr)                       # fastai_optim.py:185-190 (every trainable parameter)
        if g is None:
            continue                                   # Adam skips parameters without a gradient
        if not true_wd and wd != 0:
            g = g + wd * p
        state["step"][i] += 1
        t = state["step"][i]
        m, v = state["m"][i], state["v"][i]
        m.mul_(beta1).add_(g, alpha=1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        bc1, bc2 = 1 - beta1 ** t, 1 - beta2 ** t
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        p.addcdiv_(m, denom, value=-(lr / bc1))


def new_state(params):
    return {"step": [0] * len(params), "m": [torch.zeros_like(p) for p in params], "v": [torch.zeros_like(p) for p in params]}


def annealing_cos(start, end, pct):
    return end + (start - end) / 2 * (math.cos(math.pi * pct) + 1)


def one_cycle(step, total_step, lr_max, moms, div_factor, pct_start):
    """-> (lr, mom) the schedule sets before optimizer step `step` (LRSchedulerStep.step)."""
    a1 = int(total_step * pct_start)
    low = lr_max / div_factor
    lr = annealing_cos(low, lr_max, step / a1) if a1 > 0 else None
    mom = annealing_cos(moms[0], moms[1], step / a1) if a1 > 0 else None
    if step >= a1:
        pct = (step - a1) / (total_step - a1)
        lr, mom = annealing_cos(lr_max, low / 1e4, pct), annealing_cos(moms[1], moms[0], pct)
    return lr, mom
