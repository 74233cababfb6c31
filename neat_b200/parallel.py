"""Data parallelism over rays/images: one process per GPU, ONE all-reduce of a flat gradient bucket per step
(SURVEY.md section 8e).  Rays are independent through the whole path, so there is no other collective."""
import torch
import torch.distributed as dist


class GradBucket:
    """All parameter gradients live in one flat fp32 buffer (p.grad are views), so the data-parallel
    reduction is a single `all_reduce(SUM)` over NCCL/NVLink followed by a scale, with no packing copies.
    `extra` reserves trailing slots for loss scalars that ride in the same message."""

    def __init__(self, params, extra=0):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n + extra, dtype=torch.float32, device=dev)
        self.n, self.extra = n, extra
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def scalars(self):
        return self.flat[self.n:]

    def all_reduce_sum(self, group=None):
        """SUM only; returns the scale (1 / world size) the consumer still has to apply (neat_b200.optim.Adam folds it
        into its single update kernel: grad_scale)."""
        if not (dist.is_available() and dist.is_initialized()):
            return 1.0
        ws = dist.get_world_size(group)
        if ws == 1:
            return 1.0
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        return 1.0 / ws

    def all_reduce_mean(self, group=None):
        if not (dist.is_available() and dist.is_initialized()):
            return
        ws = dist.get_world_size(group)
        if ws == 1:
            return
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        self.flat.mul_(1.0 / ws)
