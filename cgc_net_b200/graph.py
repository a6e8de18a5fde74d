"""Host-side graph containers and the synthetic cell-graph generator.

The reference feeds ``SoftPoolingGcnEncoder.forward`` either a PyG ``Batch`` (attributes
``x``, ``edge_index``, ``batch``, ``y``; ``model/network.py:237-243``) or a dense tuple
(``network.py:254-256``).  ``Data`` / ``Batch`` below are the minimal attribute bags that
carry the same fields (``Batch.from_data_list`` follows PyG 1.2.1 ``data/batch.py``:
tensors concatenated along dim 0, ``edge_index`` along dim 1 with a running node offset,
``batch`` = graph id per node, sorted).  They additionally carry HOST-side metadata
(``num_graphs``, ``num_nodes_list``) so the CUDA path needs no device->host sync to size
its buffers (the reference pays two ``.item()`` syncs, ``model/utils.py:17,21``).

Synthetic inputs follow SURVEY.md section 8(d): nodes uniformly placed in a square,
k-nearest-neighbour edges including the self loop, ``x ~ N(0, 1)`` (the reference
z-scores its features, ``dataflow/data.py:353``), labels ``U{0,1,2}``.
"""
import torch


class Data(object):
    """One cell graph: ``x [N,F]``, ``edge_index [2,E]`` (int64), ``y [1]``, optional ``pos``."""

    def __init__(self, x=None, edge_index=None, y=None, pos=None, patch_idx=None):
        self.x, self.edge_index, self.y, self.pos, self.patch_idx = x, edge_index, y, pos, patch_idx

    @property
    def num_nodes(self):
        return self.x.shape[0]


class Batch(object):
    """Concatenation of ``Data`` objects (PyG ``Batch`` fields used by network.py:239-240)."""

    def __init__(self, x, edge_index, batch, y=None, num_nodes_list=None, max_degree=None, self_loops=None):
        self.x, self.edge_index, self.batch, self.y = x, edge_index, batch, y
        self.num_nodes_list = None if num_nodes_list is None else [int(n) for n in num_nodes_list]
        # host-side facts about the adjacency, so that the index-table pitch can be chosen without a device sync:
        # largest number of COO entries of one source row, and whether every node carries a self loop
        self.max_degree = None if max_degree is None else int(max_degree)
        self.self_loops = None if self_loops is None else bool(self_loops)

    @property
    def num_graphs(self):
        if self.num_nodes_list is not None:
            return len(self.num_nodes_list)
        return int(self.batch[-1]) + 1

    @staticmethod
    def from_data_list(data_list):
        xs, eis, bs, ys, counts = [], [], [], [], []
        offset = 0
        max_degree, self_loops = 0, True
        for g, d in enumerate(data_list):
            n = d.x.shape[0]
            xs.append(d.x)
            if max_degree is not None and not d.edge_index.is_cuda:
                if d.edge_index.shape[1]:
                    max_degree = max(max_degree, int(torch.bincount(d.edge_index[0], minlength=1).max()))
                loops = d.edge_index[0][d.edge_index[0] == d.edge_index[1]]
                self_loops = self_loops and int(torch.unique(loops).numel()) == n
            else:
                max_degree = self_loops = None          # device-resident edge lists: not inspected (no sync)
            eis.append(d.edge_index + offset)
            bs.append(torch.full((n,), g, dtype=torch.long))
            if d.y is not None:
                ys.append(d.y.view(-1))
            counts.append(n)
            offset += n
        y = torch.cat(ys) if ys else None
        return Batch(torch.cat(xs), torch.cat(eis, dim=1), torch.cat(bs), y, counts, max_degree, self_loops)

    def _map(self, fn):
        y = None if self.y is None else fn(self.y)
        return Batch(fn(self.x), fn(self.edge_index), fn(self.batch), y, self.num_nodes_list, self.max_degree,
                     self.self_loops)

    def to(self, device, non_blocking=False):
        return self._map(lambda t: t.to(device, non_blocking=non_blocking))

    def pin_memory(self):
        return self._map(lambda t: t.pin_memory())

    def nbytes(self):
        ts = [self.x, self.edge_index, self.batch] + ([] if self.y is None else [self.y])
        return sum(t.numel() * t.element_size() for t in ts)


def knn_edges(pos, k, row_is_centre=True):
    """k nearest neighbours INCLUDING the node itself, as a COO ``edge_index [2, N*k]``.

    ``row_is_centre=True``: ``edge_index[0]`` = centre, ``edge_index[1]`` = its k nearest, so
    every adjacency ROW has exactly ``k`` non-zeros (the fixed-degree roofline variant of
    SURVEY 8(d)).  ``False`` mimics ``radius_graph`` (``dataflow/data.py:348``):
    ``edge_index[0]`` = neighbour, ``edge_index[1]`` = centre (column degree fixed, row degree varies).
    """
    n = pos.shape[0]
    k = min(k, n)
    d = torch.cdist(pos, pos)
    d.fill_diagonal_(-1.0)                       # self first
    nbr = d.topk(k, dim=1, largest=False).indices   # [N, k]
    centre = torch.arange(n).view(-1, 1).expand(n, k)
    if row_is_centre:
        return torch.stack([centre.reshape(-1), nbr.reshape(-1)])
    return torch.stack([nbr.reshape(-1), centre.reshape(-1)])


def synthetic_cell_graph(num_nodes, degree, feat_dim=16, seed=0, row_is_centre=True, num_classes=3):
    """One synthetic cell graph; generator seeded with ``1000 + seed``."""
    g = torch.Generator().manual_seed(1000 + int(seed))
    pos = torch.rand(num_nodes, 2, generator=g) * 1000.0
    x = torch.randn(num_nodes, feat_dim, generator=g)
    y = torch.randint(0, num_classes, (1,), generator=g)
    return Data(x=x, edge_index=knn_edges(pos, degree, row_is_centre), y=y, pos=pos)


def synthetic_batch(num_graphs, num_nodes=1800, degree=8, feat_dim=16, seed=0, row_is_centre=True,
                    ragged=None):
    """Batch of synthetic graphs.  ``ragged=(nmin, nmax, dmin, dmax)`` draws size and degree
    per graph (BASELINE config #5: 500-3000 nodes, degree 5-12)."""
    g = torch.Generator().manual_seed(77 + int(seed))
    graphs = []
    for i in range(num_graphs):
        n, k = num_nodes, degree
        if ragged is not None:
            n = int(torch.randint(ragged[0], ragged[1] + 1, (1,), generator=g))
            k = int(torch.randint(ragged[2], ragged[3] + 1, (1,), generator=g))
        graphs.append(synthetic_cell_graph(n, k, feat_dim, seed * 100003 + i, row_is_centre))
    return graphs


def split_by_node_count(data_list, world_size):
    """Contiguous split with ~equal cumulative node count per rank -- PyG 1.2.1
    ``DataParallel.scatter`` (``nn/data_parallel.py``; call site train.py:280,285):
    ``device_id = floor(R * midpoint_cumsum / total)``."""
    counts = torch.tensor([d.num_nodes for d in data_list], dtype=torch.float64)
    cum = counts.cumsum(0)
    mid = (cum - counts / 2.0) * world_size / cum[-1]
    dev = mid.floor().long().clamp_(max=world_size - 1)
    return [[d for d, r in zip(data_list, dev.tolist()) if r == rank] for rank in range(world_size)]


def radius_graph(pos, r, batch=None, loop=False, max_num_neighbors=32, num_nodes_list=None):
    """``torch_cluster.radius_graph(pos, r, batch, loop, max_num_neighbors)`` (flow ``source_to_target``) as the
    reference's loaders call it (dataflow/data.py:246,255,297,348 with ``loop=True, max_num_neighbors=8``), on
    the GPU: ``edge_index[0]`` = neighbour, ``edge_index[1]`` = centre, <= k nearest within ``r`` per centre,
    centre-major, nearest first.  ``pos [N, 2]`` CUDA float tensor; ``batch`` sorted graph ids (or ``None`` for a
    single graph); ``num_nodes_list`` avoids the device->host sync for the sizes."""
    from . import _backend
    be = _backend.get()
    dev = pos.device
    n_total = pos.shape[0]
    if num_nodes_list is None:
        if batch is None:
            num_nodes_list = [n_total]
        else:
            num_nodes_list = torch.bincount(batch).tolist()
    counts = torch.tensor(num_nodes_list, dtype=torch.int32)
    start = torch.zeros(len(num_nodes_list) + 1, dtype=torch.int32)
    start[1:] = torch.cumsum(counts, 0)
    start = start.to(dev)
    k = int(max_num_neighbors)
    nbr = torch.empty(n_total, k, dtype=torch.int32, device=dev)
    deg = torch.empty(n_total, dtype=torch.int32, device=dev)
    be.radius_graph(pos.contiguous().float(), start, len(num_nodes_list), int(max(num_nodes_list)), r, k, loop, nbr, deg)
    ends = torch.cumsum(deg.long(), 0)
    num_edges = int(ends[-1])                       # output size is data dependent, like the reference op
    offsets = ends - deg.long()
    edge_index = torch.empty(2, num_edges, dtype=torch.long, device=dev)
    if num_edges:
        be.knn_to_coo(nbr, deg, offsets, k, num_edges, edge_index)
    return edge_index


def sample_nodes(pos, num_nodes_list, ratio, method='fuse', generator=None):
    """Node subsampling of the reference's loaders (``_sampling``, dataflow/data.py:195-225) on the GPU: per graph
    ``int(n * ratio)`` node ids by ``method``:

    * ``'random'``   -- ``np.random.choice(n, k, replace=False)`` (data.py:223): a random subset;
    * ``'farthest'`` -- ``FarthestSampler`` (common/utils.py:187-197) from a random first node: ``cgc_farthest_point_sample``;
    * ``'fuse'``     -- 70 % farthest-point, 30 % drawn at random from the rest (data.py:210-219).

    Returns a list of int64 tensors of LOCAL node ids (one per graph, farthest-point ids first).  The random
    draws come from ``generator`` (a CPU ``torch.Generator``), not from numpy's global stream."""
    from . import _backend
    be = _backend.get()
    dev = pos.device
    counts = [int(n) for n in num_nodes_list]
    ks = [int(n * ratio) for n in counts]
    if method == 'random':
        return [torch.randperm(n, generator=generator)[:k].to(dev) for n, k in zip(counts, ks)]
    far = ks if method == 'farthest' else [int(0.7 * k) for k in ks]
    start = torch.tensor([0] + counts, dtype=torch.int64).cumsum(0).to(torch.int32)
    ostart = torch.tensor([0] + far, dtype=torch.int64).cumsum(0).to(torch.int32)
    first = torch.tensor([int(torch.randint(0, max(n, 1), (1,), generator=generator)) for n in counts],
                         dtype=torch.int32)
    out = torch.zeros(max(int(ostart[-1]), 1), dtype=torch.int32, device=dev)
    if int(ostart[-1]) > 0:
        be.farthest_point_sample(pos.contiguous().float(), start.to(dev), first.to(dev), ostart.to(dev),
                                 max(counts), out)
    result = []
    os_ = ostart.tolist()
    for g, (n, k) in enumerate(zip(counts, ks)):
        ids = out[os_[g]:os_[g + 1]].long()
        if method == 'fuse' and k > far[g]:
            taken = torch.zeros(n, dtype=torch.bool)
            taken[ids.cpu()] = True
            rest = (~taken).nonzero().view(-1)
            extra = rest[torch.randperm(rest.numel(), generator=generator)[:k - far[g]]]
            ids = torch.cat([ids, extra.to(dev)])
        result.append(ids)
    return result


class DevicePrefetcher(object):
    """Wrap an iterable of HOST ``Batch`` objects (pinned memory) and yield them on ``device``, copying batch
    ``i + 1`` on a side stream while the caller computes on batch ``i`` -- the loader-side counterpart of the
    reference's ``DataListLoader(num_workers=...)`` + ``.to(device)`` inside ``DataParallel`` (train.py:174-178,
    dataflow/data.py:52), so the 23 MB host->device copy of a 64-graph batch leaves the critical path.

    Two persistent device slots (grown on demand) are recycled: no allocator traffic per batch, and a slot is
    overwritten only after the compute stream has passed the point where its previous batch was handed back."""

    def __init__(self, batches, device, slots=2):
        self.batches, self.device = batches, torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self.slots = [dict(bufs={}, free=None) for _ in range(slots)]

    def _buffer(self, slot, name, host):
        buf = slot['bufs'].get(name)
        if buf is None or buf.dtype != host.dtype or buf.numel() < host.numel():
            buf = torch.empty(max(host.numel(), 1), dtype=host.dtype, device=self.device)
            slot['bufs'][name] = buf
            slot['grown'] = True        # fresh block of the caching allocator: see _issue
        return buf[:host.numel()].view(host.shape)

    def _issue(self, host_batch, slot):
        views = {}
        for name in ('x', 'edge_index', 'batch', 'y'):
            t = getattr(host_batch, name)
            views[name] = None if t is None else self._buffer(slot, name, t)      # allocation on the caller's stream
        with torch.cuda.stream(self.stream):
            if slot.pop('grown', False):
                # a buffer was (re)allocated on the caller's stream: the caching allocator may have handed out a
                # block whose previous user still has kernels pending there, so the copy waits for that stream
                self.stream.wait_stream(torch.cuda.current_stream(self.device))
            if slot['free'] is not None:
                self.stream.wait_event(slot['free'])           # the slot's previous batch has been consumed
            for name, dst in views.items():
                if dst is not None:
                    dst.copy_(getattr(host_batch, name), non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.stream)
        return Batch(views['x'], views['edge_index'], views['batch'], views['y'], host_batch.num_nodes_list,
                     getattr(host_batch, 'max_degree', None), getattr(host_batch, 'self_loops', None)), done

    def __iter__(self):
        it = iter(self.batches)
        compute = torch.cuda.current_stream(self.device)
        n = len(self.slots)
        i = 0
        try:
            nxt = self._issue(next(it), self.slots[0])
        except StopIteration:
            return
        while nxt is not None:
            cur, done = nxt
            i += 1
            try:
                nxt = self._issue(next(it), self.slots[i % n])
            except StopIteration:
                nxt = None
            compute.wait_event(done)
            yield cur
            # the caller is back: everything it enqueued on the compute stream so far used `cur`
            free = torch.cuda.Event()
            free.record(compute)
            self.slots[(i - 1) % n]['free'] = free
