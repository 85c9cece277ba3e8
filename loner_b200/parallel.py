"""Ray-sharded data parallelism: host-side logic of the two exchanges per step.

Every rank holds all keyframes and the full (replicated) parameters / optimiser state; rank r draws
its own rays.  Exactness w.r.t. the single-GPU loss needs GLOBAL normalisers, so the step does
  (1) all-reduce of the two loss normalisers (#valid, #opaque rays)   -- before the loss kernel
  (2) all-reduce of ONE flat buffer  [MLP grads | pose grads | 4 loss sums]  -- after backward
The same code runs over NCCL (GPU box) and gloo (CPU tests).
"""
import torch
import torch.distributed as dist


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard_slice(n_total: int, r: int, world: int):
    """Contiguous slice [lo, hi) of n_total rays owned by rank r (remainder spread over the first ranks)."""
    base, rem = divmod(n_total, world)
    lo = r * base + min(r, rem)
    return lo, lo + base + (1 if r < rem else 0)


def allreduce_counts(counters: torch.Tensor):
    """int32[2] (#valid, #opaque) -> global sums, in place."""
    if world_size() > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    return counters


class FlatExchange:
    """Packs several gradient tensors into one flat fp32 buffer for a single all-reduce."""

    def __init__(self, shapes, device):
        self.shapes = [tuple(s) for s in shapes]
        self.sizes = [int(torch.tensor(s).prod()) if len(s) else 1 for s in self.shapes]
        self.flat = torch.zeros(sum(self.sizes), device=device, dtype=torch.float32)

    def views(self):
        out, off = [], 0
        for s, n in zip(self.shapes, self.sizes):
            out.append(self.flat[off:off + n].view(s))
            off += n
        return out

    def reduce(self, tensors):
        """Sums `tensors` (list matching shapes) over ranks; returns views into the flat buffer."""
        views = self.views()
        for v, t in zip(views, tensors):
            v.copy_(t)
        if world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        return views
