"""clip_grad_norm_ + OptimWrapper.step() as two kernel launches (csrc/optim.cu).

Reference semantics restated (cited lines are /root/reference):
  * `torch.nn.utils.clip_grad_norm_(net.parameters(), 10.0)`                               train_hdf5.py:671
  * `OptimWrapper.step()`: with `true_wd` every trainable parameter of every layer group - BatchNorm ones too,
    `bn_wd=True` - is multiplied by `1 - wd*lr`, Adam's own weight_decay is forced to 0, then `opt.step()`
                                                                                rslo/torchplus/train/fastai_optim.py:181-194
  * the inner optimizer is `torch.optim.Adam(betas=(0.9, 0.99))` whose beta1 the schedule overwrites through `.mom`
    every step; one lr for all layer groups                              rslo/builder/optimizer_builder.py:101-118
  * `OneCycle`: cosine lr_max/div -> lr_max over pct_start, then -> lr_max/div/1e4; momentum moms[0] -> moms[1] -> moms[0]
                                                               rslo/torchplus/train/learning_schedules_fastai.py:64-95
  * parameters whose `.grad` is None are skipped by Adam (no moment update, no step count) but still decayed.

B200 design: the gradients are already slices of one flat buffer (`FlatGradAllReducer`, the NCCL bucket); the two
Adam moments are flat buffers with the same offsets; `rslo_grad_sumsq` reduces the global norm into a device double
(bit-reproducible), `rslo_adam_step` applies clip coefficient, 1/world averaging, decay and the Adam update in one
streaming pass - no host synchronisation anywhere (the reference's clip does `.item()`-free but ~900 small launches).
Parameters stay in their own storage (state_dict layout untouched); in-place writes go through `.data` pointers, so
weight-image caches keyed on tensor versions must be told: `step()` bumps every parameter's version counter.
"""
import ctypes as C
import math

import numpy as np
import torch

from ... import kernels as K
from ..._lib import AdamChunk
from ...layers.sparse3d import invalidate_weight_images

ADAM_CHUNK = 8192


def annealing_cos(start, end, pct):
    """`learning_schedules_fastai.py:57-61`"""
    return end + (start - end) / 2 * (math.cos(math.pi * pct) + 1)


class OneCycle:
    """`learning_schedules_fastai.py:64-95` + `LRSchedulerStep.step` (:44-52): sets optimizer.lr / .mom for `step`."""

    def __init__(self, optimizer, total_step, lr_max, moms, div_factor, pct_start):
        self.optimizer, self.total_step = optimizer, total_step
        self.lr_max, self.moms, self.div_factor, self.pct_start = lr_max, list(moms), div_factor, pct_start
        low = lr_max / div_factor
        a1 = int(total_step * pct_start)
        self.lr_phases = [(0, a1, (low, lr_max)), (a1, total_step, (lr_max, low / 1e4))]
        self.mom_phases = [(0, a1, (self.moms[0], self.moms[1])), (a1, total_step, (self.moms[1], self.moms[0]))]
        optimizer.lr, optimizer.mom = low, self.moms[0]

    @staticmethod
    def _value(phases, step):
        val = None
        for start, end, (a, b) in phases:
            if step >= start:
                val = annealing_cos(a, b, (step - start) / (end - start))
        return val

    def step(self, step):
        self.optimizer.lr = self._value(self.lr_phases, step)
        self.optimizer.mom = self._value(self.mom_phases, step)

    @property
    def learning_rate(self):
        return self.optimizer.lr


class FusedAdamClip:
    """Drop-in for the pair (clip_grad_norm_, OptimWrapper over Adam) on one device.

    reducer: a `FlatGradAllReducer` (its `flat` buffer and parameter order are used); call after
    `reducer.all_reduce(average=False)` or `reducer.pack()` - the 1/world averaging is folded into the step."""

    def __init__(self, reducer, lr=3e-3, wd=0.0, true_wd=True, bn_wd=True, betas=(0.9, 0.99), eps=1e-8, max_norm=10.0,
                 write_clipped_grad=False):
        assert bn_wd, "bn_wd=False is outside the shipped configs (optimizer_builder.py:117)"
        self.reducer = reducer
        self.params = reducer.params
        self.flat = reducer.flat
        assert self.flat.dtype == torch.float32 and self.flat.is_cuda and self.flat.numel() < 2 ** 32
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.lr, self.mom, self.beta, self.wd = float(lr), float(betas[0]), float(betas[1]), float(wd)
        self.true_wd, self.eps, self.max_norm = bool(true_wd), float(eps), float(max_norm)
        self.write_clipped_grad = write_clipped_grad
        self.step_count = 0
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=self.flat.device)
        self._ws = torch.zeros(K.grad_norm_workspace_bytes(), dtype=torch.uint8, device=self.flat.device)
        self._tables = {}
        self._sig = tuple(p.data_ptr() for p in self.params)
        self.name = "adam_optimizer"

    # ---- chunk table: one entry per <= ADAM_CHUNK contiguous elements of one parameter ---------------------------
    def _table(self, present):
        key = (tuple(p.data_ptr() for p in self.params), present)
        tab = self._tables.get(key)
        if tab is None:
            rows = []
            off = 0
            for p, has in zip(self.params, present):
                assert p.is_contiguous()
                n, base = p.numel(), p.data_ptr()
                for c in range(0, n, ADAM_CHUNK):
                    rows.append((base + 4 * c, off + c, min(ADAM_CHUNK, n - c), 1 if has else 0))
                off += n
            arr = (AdamChunk * len(rows))()
            for i, (ptr, o, n, f) in enumerate(rows):
                arr[i].p, arr[i].off, arr[i].n, arr[i].flags = ptr, o, n, f
            host = torch.from_numpy(np.frombuffer(bytes(arr), dtype=np.uint8).copy())
            tab = (host.to(self.flat.device), len(rows))
            if len(self._tables) > 8:
                self._tables.clear()
            self._tables[key] = tab
        return tab

    # ---- the reference's call sequence ---------------------------------------------------------------------------
    def zero_grad(self):
        self.reducer.zero_()

    def clip_grad_norm_(self, max_norm=None):
        """Computes the global norm on the device; the clip itself happens inside step().  Returns the norm of the
        averaged gradient as a 0-d device tensor (no synchronisation)."""
        if max_norm is not None:
            self.max_norm = float(max_norm)
        K.grad_sumsq(self.flat, self.sumsq, self._ws)
        self._norm_fresh = True
        return self.sumsq.sqrt().to(torch.float32) * self._grad_scale()

    def _grad_scale(self):
        return 1.0 / self.reducer.world if getattr(self.reducer, "pending_average", False) else 1.0

    def step(self):
        self.reducer.pack()                        # no-op when all_reduce() already packed this step's gradients
        present = tuple(self.reducer.present)
        tab, n = self._table(present)
        if self.max_norm > 0 and not getattr(self, "_norm_fresh", False):
            K.grad_sumsq(self.flat, self.sumsq, self._ws)
        self._norm_fresh = False
        self.step_count += 1
        K.adam_step(tab, n, self.flat, self.exp_avg, self.exp_avg_sq, self.sumsq if self.max_norm > 0 else None,
                    self._grad_scale(), self.max_norm, self.lr, self.mom, self.beta, self.eps, self.wd, self.true_wd,
                    self.step_count, self.write_clipped_grad)
        if getattr(self.reducer, "pending_average", False):
            self.reducer.pending_average = False
        invalidate_weight_images()                 # the kernel wrote through raw pointers: version counters did not move

    # ---- checkpointing (by parameter order of the reducer) -------------------------------------------------------
    def state_dict(self):
        return {"step": self.step_count, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "lr": self.lr,
                "mom": self.mom, "beta": self.beta, "wd": self.wd}

    def load_state_dict(self, sd):
        self.step_count = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.lr, self.mom, self.beta, self.wd = float(sd["lr"]), float(sd["mom"]), float(sd["beta"]), float(sd["wd"])


def build_optimizer(optimizer_config, reducer, total_step=None):
    """`optimizer_builder.build` + `lr_scheduler_builder.build` for the shipped `adam_optimizer { one_cycle }` config:
    -> (FusedAdamClip, OneCycle or None).  optimizer_config: the parsed `train_config.optimizer` message."""
    cfg = optimizer_config.adam_optimizer
    fixed = bool(getattr(optimizer_config, "fixed_weight_decay", False))
    opt = FusedAdamClip(reducer, lr=3e-3, wd=float(cfg.weight_decay), true_wd=fixed, bn_wd=True,
                        betas=(0.9, 0.99) if fixed else (0.9, 0.999), eps=1e-8, max_norm=10.0)
    sched = None
    lr_cfg = getattr(cfg, "learning_rate", None)
    oc = getattr(lr_cfg, "one_cycle", None) if lr_cfg is not None else None
    if oc is not None and total_step:
        sched = OneCycle(opt, total_step, float(oc.lr_max), list(oc.moms), float(oc.div_factor), float(oc.pct_start))
    return opt, sched
