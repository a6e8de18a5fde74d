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
        self.views, self.offsets = [], []
        off = 0
        for p in self.params:
            self.views.append(self.flat[off: off + p.numel()].view_as(p))
            self.offsets.append(off)
            off += p.numel()
        self._overlap = None

    def zero(self):
        """Use instead of ``optimizer.zero_grad()``."""
        for p in self.params:
            p.grad = None
        if self._overlap is not None:
            self._overlap.reset()

    def enable_overlap(self, chunks=2, group=None):
        """Start the all-reduce of a chunk of the bucket as soon as the backward pass has produced every gradient in
        it, so the collective runs under the rest of the backward.  The parameter list is cut
        into ``chunks`` contiguous ranges of roughly equal BYTES; autograd reaches parameters roughly in reverse
        order, so the last range (read-out head, stage 3, stage 2) is reduced while stage 1 is still differentiating.
        ``all_reduce_mean()`` then only launches what is left (parameters that received no gradient), waits for the
        handles and applies 1/R.  Call once after construction; a no-op without an initialised process group."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return False
        self._overlap = _EarlyReduce(self, max(1, int(chunks)), group)
        return True

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
        if self._overlap is not None:
            self._overlap.finish()
            self.flat.mul_(1.0 / world)
            return
        self.pack()
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        self.flat.mul_(1.0 / world)


class _EarlyReduce(object):
    """Per-chunk asynchronous all-reduce driven by post-accumulate-grad hooks (see FlatGradBucket.enable_overlap)."""

    def __init__(self, bucket, chunks, group):
        self.bucket, self.group = bucket, group
        n = len(bucket.params)
        total = bucket.flat.numel()
        # contiguous parameter ranges of ~equal bytes
        bounds, acc, target = [0], 0, total / float(chunks)
        for i, p in enumerate(bucket.params):
            acc += p.numel()
            if acc >= target * len(bounds) and len(bounds) < chunks and i + 1 < n:
                bounds.append(i + 1)
        bounds.append(n)
        self.ranges = [(bounds[c], bounds[c + 1]) for c in range(len(bounds) - 1)]
        self.chunk_of = {}
        for c, (lo, hi) in enumerate(self.ranges):
            for i in range(lo, hi):
                self.chunk_of[i] = c
        self.handles = []
        for i, p in enumerate(bucket.params):
            p.register_post_accumulate_grad_hook(self._make_hook(i))
        self.reset()

    def reset(self):
        self.pending = [hi - lo for lo, hi in self.ranges]
        self.seen = [False] * len(self.bucket.params)
        self.launched = [False] * len(self.ranges)
        self.handles = []

    def _span(self, c):
        lo, hi = self.ranges[c]
        b = self.bucket
        end = b.offsets[hi] if hi < len(b.params) else b.flat.numel()
        return b.flat[b.offsets[lo]: end]

    def _launch(self, c):
        self.launched[c] = True
        self.handles.append(dist.all_reduce(self._span(c), op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def _make_hook(self, i):
        def hook(p):
            b = self.bucket
            v = b.views[i]
            if p.grad is not None and (p.grad.data_ptr() != v.data_ptr() or p.grad.shape != v.shape):
                v.copy_(p.grad)
                p.grad = v
            if not self.seen[i]:
                self.seen[i] = True
                c = self.chunk_of[i]
                self.pending[c] -= 1
                if self.pending[c] == 0 and not self.launched[c]:
                    self._launch(c)
        return hook

    def finish(self):
        b = self.bucket
        for i, p in enumerate(b.params):                 # parameters autograd never reached: zero gradient
            if not self.seen[i]:
                if p.grad is None:
                    b.views[i].zero_()
                elif p.grad.data_ptr() != b.views[i].data_ptr():
                    b.views[i].copy_(p.grad)
                p.grad = b.views[i]
        for c in range(len(self.ranges)):
            if not self.launched[c]:
                self._launch(c)
        for h in self.handles:
            h.wait()
        self.reset()


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
