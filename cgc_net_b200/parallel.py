"""Data parallelism for the hot path: one process per GPU, one gradient all-reduce per step.

The reference uses single-process ``DataParallel`` (``train.py:276-287``): parameters are
re-broadcast every iteration, per-replica mean losses are averaged (``train.py:179``) and
gradients are reduced onto device 0.  Here every rank owns a replica; graphs are sharded by
batch index (``graph.split_by_node_count`` mirrors PyG's node-balanced scatter), BatchNorm stays
per-replica as in the reference, and the only collective is ONE ``all_reduce(sum)`` over a flat
fp32 bucket holding every gradient (1 410 275 elements = 5.64 MB for the BASELINE model),
followed by a scale of ``1 / world_size`` -- NCCL over NVLink/NVSwitch on B200, gloo in the CPU
tests.  No data-path collective exists: graphs never interact across ranks.
"""
import torch
import torch.distributed as dist


class FlatGradBucket(object):
    """One contiguous buffer for every parameter gradient, filled once per step.

    ``zero()`` drops the ``.grad`` tensors, so autograd *assigns* each freshly computed gradient instead of
    launching one ``grad += new`` kernel per parameter (60 launches per step for the BASELINE model).
    ``all_reduce_mean()`` packs them into the flat buffer with one multi-tensor copy, all-reduces it once, and
    leaves every ``.grad`` as a view of the buffer (what the optimiser then reads).  With a single rank nothing
    is copied at all."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(total, dtype=ref.dtype, device=ref.device)
        self.views = []
        off = 0
        for p in self.params:
            self.views.append(self.flat[off: off + p.numel()].view_as(p))
            off += p.numel()

    def zero(self):
        """Use instead of ``optimizer.zero_grad()``."""
        for p in self.params:
            p.grad = None

    def nbytes(self):
        return self.flat.numel() * self.flat.element_size()

    def pack(self):
        """Copy the per-parameter gradients into the flat buffer (missing gradients count as zero) and make
        every ``.grad`` a view of it.  Idempotent: gradients that already ARE the views (a second ``pack()``,
        ``pack()`` followed by ``all_reduce_mean()``, gradient accumulation without ``zero()``) stay untouched."""
        have = []
        for v, p in zip(self.views, self.params):
            if p.grad is None:
                v.zero_()                       # only a MISSING gradient is zeroed; a live view keeps its content
            elif p.grad.data_ptr() != v.data_ptr() or p.grad.shape != v.shape:
                have.append((v, p.grad))
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        for p, v in zip(self.params, self.views):
            p.grad = v

    def all_reduce_mean(self, group=None):
        """Sum over ranks, then 1/R: the mean of per-replica mean losses (train.py:179)."""
        if not (dist.is_available() and dist.is_initialized()):
            return
        world = dist.get_world_size(group)
        if world == 1:
            return
        self.pack()
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        self.flat.mul_(1.0 / world)


def broadcast_parameters(module, src=0, group=None):
    """Once at start-up (the reference re-broadcasts every iteration inside DataParallel.replicate)."""
    if not (dist.is_available() and dist.is_initialized()):
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)


def shard_for_rank(data_list, rank, world_size, by_nodes=False):
    """Contiguous shard of the list of graphs for ``rank``.  ``by_nodes`` = PyG's node-balanced split."""
    if by_nodes:
        from .graph import split_by_node_count
        return split_by_node_count(data_list, world_size)[rank]
    per = (len(data_list) + world_size - 1) // world_size
    return data_list[rank * per: (rank + 1) * per]
