"""Data-parallel plumbing for the hot path: frame pairs shard across ranks, the only exchange is one
all-reduce of the gradients (reference: apex DistributedDataParallel, `train_hdf5.py:463`, 12.0 M fp32
= 48 MB per step; or `average_gradients`, `rslo/utils/distributed_utils.py:53-65`, one call per tensor).

B200 design: the gradients are packed into ONE flat fp32 buffer by multi-tensor copies, so the step's exchange
is two in-place NCCL all-reduces over NVLink/NVSwitch (no per-tensor calls; ~0.15 ms at wire speed for 48 MB on
8 ranks): the dense head's bucket (44 MB) goes out on a communication stream as soon as the head's backward is
done, overlapping the sparse encoder's backward; the rest follows right after backward.
"""
import datetime
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """One process per GPU, rendezvous from RANK / WORLD_SIZE / MASTER_* (torchrun)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local),
                                    timeout=datetime.timedelta(seconds=180))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world, timeout=datetime.timedelta(seconds=180))
    return rank, local, world


class FlatGradAllReducer:
    """Owns one flat gradient buffer whose slices become every trainable parameter's `.grad`.

    Cycle per step: `zero_()` drops the gradients (so the first backward of the step STEALS its gradients
    instead of launching one `grad += new` kernel per parameter), any number of backward passes, then
    `all_reduce()` packs whatever gradients exist into the flat buffer with one multi-tensor copy, reduces
    it in place over NCCL and leaves `p.grad` pointing at the parameter's slice (zeros where a parameter
    received no gradient — ~0.74 M parameters of the shipped head never do).

    Overlap: the flat buffer is laid out [early | rest]; `early` = the parameters of `early_module` (the dense head,
    whose backward finishes first).  When the module reports that its backward is done (`net.on_head_backward_done`
    fires from an autograd hook on the head's input), the early bucket is packed and its all-reduce is launched on a
    communication stream while the sparse encoder's backward is still running; `all_reduce()` then only has the
    remaining bucket to pack and reduce, plus one in-place 1/world scaling of the flat buffer."""

    def __init__(self, module, process_group=None, early_module=None):
        params = [p for p in module.parameters() if p.requires_grad]
        early_ids = set()
        if early_module is not None:
            early_ids = {id(p) for p in early_module.parameters() if p.requires_grad}
        self.early = [p for p in params if id(p) in early_ids]
        self.rest = [p for p in params if id(p) not in early_ids]
        self.params = self.early + self.rest
        n = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(n, dtype=ref.dtype, device=ref.device)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.n_early = sum(p.numel() for p in self.early)
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self._comm = None
        self._early_work = None
        self._early_done = False
        self.present = [False] * len(self.params)       # which parameters received a gradient (recorded by pack)
        self._packed = False
        if early_module is not None and self.world > 1 and ref.is_cuda and hasattr(module, "on_head_backward_done"):
            self._comm = torch.cuda.Stream(device=ref.device)
            module.on_head_backward_done = self._reduce_early

    @property
    def nbytes(self):
        return self.flat.numel() * self.flat.element_size()

    def zero_(self):
        for p in self.params:
            p.grad = None
        self._early_done = False
        self._early_work = None
        self._packed = False
        self.pending_average = False

    def broadcast_params(self, src=0):
        """Same replica everywhere: parameters AND buffers (BatchNorm running statistics, global_step)
        (reference: `broadcast_params`, distributed_utils.py:68-71).  Writes go through detached views under no_grad
        so tensor version counters advance (weight-image caches key on them)."""
        if self.world > 1:
            with torch.no_grad():
                for p in self.params:
                    dist.broadcast(p.detach(), src, group=self.group)

    def sync_buffers(self, module, src=None):
        """BatchNorm statistics stay per rank during training (north_star: gradients only; the reference's apex SyncBN
        also reduces them).  Call this before checkpointing / evaluation: floating-point buffers are averaged over the
        ranks (or broadcast from `src`), integer buffers broadcast from rank 0, so every rank holds the same state_dict."""
        if self.world == 1:
            return
        with torch.no_grad():
            for b in module.buffers():
                if src is not None or not b.is_floating_point():
                    dist.broadcast(b, 0 if src is None else src, group=self.group)
                else:
                    dist.all_reduce(b, op=dist.ReduceOp.SUM, group=self.group)
                    b.div_(self.world)

    def _pack(self, params, views, first=0):
        """Gradients -> flat buffer slices (one multi-tensor copy); `p.grad` becomes the slice.  Slices of parameters
        without a gradient are zeroed individually (gradients that already live in their slices stay untouched)."""
        src, dst, missing = [], [], []
        for i, (p, v) in enumerate(zip(params, views)):
            if p.grad is None:
                missing.append(v)
                self.present[first + i] = False
                continue
            self.present[first + i] = True
            if p.grad.data_ptr() != v.data_ptr():
                src.append(p.grad)
                dst.append(v)
        if missing:
            torch._foreach_zero_(missing)
        if src:
            torch._foreach_copy_(dst, src)
        for p, v in zip(params, views):
            p.grad = v

    def pack(self):
        ne = len(self.early)
        if self._packed:            # a second pack of the same step would mark the zero-filled slices as gradients
            return self.flat
        if not self._early_done:
            self._pack(self.early, self.views[:ne])
        self._pack(self.rest, self.views[ne:], ne)
        self._packed = True
        return self.flat

    def _reduce_early(self):
        """autograd-hook callback: the head's backward is done -> pack + all-reduce its bucket on the comm stream"""
        if self._early_done or self.world == 1:
            return
        ne = len(self.early)
        cur = torch.cuda.current_stream()
        self._comm.wait_stream(cur)
        with torch.cuda.stream(self._comm):
            self._pack(self.early, self.views[:ne])
            bucket = self.flat[:self.n_early]
            self._early_work = dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._early_done = True

    pending_average = False

    def all_reduce(self, average=True):
        """sum -> average over ranks, in place; one NCCL call per bucket (the early bucket may already be in flight).
        average=False leaves the SUM in the buffer and sets `pending_average`: the fused optimizer step
        (torchplus/train/fused_optim.py) folds the 1/world factor into its own pass over the gradients."""
        self.pack()
        if self.world > 1:
            if self._early_done:
                rest = self.flat[self.n_early:]
                dist.all_reduce(rest, op=dist.ReduceOp.SUM, group=self.group)
                self._early_work.wait()
                torch.cuda.current_stream().wait_stream(self._comm)
                self._early_work = None
            else:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            if average:
                self.flat.div_(self.world)
            else:
                self.pending_average = True
        return self.flat
