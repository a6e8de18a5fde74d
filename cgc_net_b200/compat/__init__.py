"""Host-side stand-ins for the third-party packages CGC-Net's ``train.py`` / ``dataflow`` import, so that the
reference training script runs UNCHANGED on top of the B200 path in an environment without
``torch_geometric==1.2.1`` / ``torch_cluster`` / ``tensorboardX``.

``install()`` registers, in ``sys.modules``:

* ``torch_geometric.data``  -- ``Data``, ``Batch``, ``Dataset``, ``DataLoader``, ``DataListLoader``
  (reference imports: dataflow/data.py:4-5);
* ``torch_geometric.nn``    -- ``DataParallel`` (train.py:17,280,285), ``radius_graph`` (dataflow/data.py:6,
  common/utils.py:4), and the parameter-holder convs of ``cgc_net_b200.network``;
* ``torch_geometric.utils`` -- ``sparse_to_dense``, ``dense_to_sparse``, ``to_dense_batch``, ``scatter_``
  (dataflow/data.py:11, common/utils.py:3, model/network.py:5, model/utils.py:2);
* ``tensorboardX.SummaryWriter`` (train.py:6) as a no-op logger;
* optionally ``model`` / ``model.network`` -> ``cgc_net_b200.network``, which is the whole drop-in: the
  reference's ``from model import network`` (train.py:16) then builds the B200 encoder.

Semantics follow the pinned versions (requirements.txt:38-42) as restated in SURVEY.md Appendix A.  This is
loader / control-plane glue in plain Python; the numerical path stays in ``cgc_net_b200.ops``.  Nothing here
imports ``oracle/``.
"""
import sys
import types

import torch

from .. import graph as _graph


# --------------------------------------------------------------------------------------- torch_geometric.data
class Data(object):
    """``torch_geometric.data.Data`` @1.2.1: attribute bag of tensors; iterating yields ``(key, item)`` pairs
    (the reference filters node-level tensors that way, dataflow/data.py:241-243,340-342)."""

    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, pos=None, **kwargs):
        self.x, self.edge_index, self.edge_attr, self.y, self.pos = x, edge_index, edge_attr, y, pos
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def keys(self):
        return [k for k, v in self.__dict__.items() if v is not None and not k.startswith('_')]

    def __getitem__(self, key):
        return getattr(self, key)

    def __setitem__(self, key, value):
        setattr(self, key, value)

    def __iter__(self):
        for k in sorted(self.keys):
            yield k, getattr(self, k)

    def __contains__(self, key):
        return key in self.keys

    @property
    def num_nodes(self):
        for k in ('x', 'pos'):
            v = getattr(self, k, None)
            if v is not None:
                return v.size(0)
        return None

    def to(self, device, **kw):
        for k in self.keys:
            v = getattr(self, k)
            if torch.is_tensor(v):
                setattr(self, k, v.to(device, **kw))
        return self


class Batch(_graph.Batch):
    """``Batch.from_data_list`` @1.2.1 (data/batch.py): tensors concatenated along dim 0, ``edge_index`` along
    dim 1 with a running node offset, ``batch`` = graph id per node.  Extra per-graph attributes the reference
    attaches (``pos``, ``patch_idx``) are concatenated too."""

    @staticmethod
    def from_data_list(data_list):
        base = _graph.Batch.from_data_list(data_list)
        out = Batch(base.x, base.edge_index, base.batch, base.y, base.num_nodes_list, base.max_degree, base.self_loops)
        for key in ('pos', 'patch_idx'):
            vals = [getattr(d, key, None) for d in data_list]
            if all(torch.is_tensor(v) for v in vals):
                setattr(out, key, torch.cat([v.view(-1) if v.dim() == 0 else v for v in vals], dim=0))
        return out

    def _map(self, fn):
        m = _graph.Batch._map(self, fn)
        out = Batch(m.x, m.edge_index, m.batch, m.y, m.num_nodes_list, m.max_degree, m.self_loops)
        for key in ('pos', 'patch_idx'):
            if hasattr(self, key):
                setattr(out, key, fn(getattr(self, key)))
        return out


class Dataset(torch.utils.data.Dataset):
    """``torch_geometric.data.Dataset`` @1.2.1: ctor runs ``_download()`` / ``_process()`` (the reference
    overrides both with no-ops, dataflow/data.py:182-193); ``__getitem__`` = ``get(idx)`` + transform."""

    def __init__(self, root=None, transform=None, pre_transform=None, pre_filter=None):
        super(Dataset, self).__init__()
        self.root, self.transform, self.pre_transform, self.pre_filter = root, transform, pre_transform, pre_filter
        self._download()
        self._process()

    def _download(self):
        pass

    def _process(self):
        pass

    def get(self, idx):
        raise NotImplementedError

    def __getitem__(self, idx):
        data = self.get(idx)
        return data if self.transform is None else self.transform(data)


class DataLoader(torch.utils.data.DataLoader):
    """Collates a list of ``Data`` into one ``Batch``."""

    def __init__(self, dataset, batch_size=1, shuffle=False, **kwargs):
        super(DataLoader, self).__init__(dataset, batch_size, shuffle,
                                         collate_fn=lambda data_list: Batch.from_data_list(data_list), **kwargs)


class DataListLoader(torch.utils.data.DataLoader):
    """Returns the raw list of ``Data`` (consumed by ``torch_geometric.nn.DataParallel``; dataflow/data.py:52,81)."""

    def __init__(self, dataset, batch_size=1, shuffle=False, **kwargs):
        super(DataListLoader, self).__init__(dataset, batch_size, shuffle, collate_fn=lambda data_list: data_list,
                                             **kwargs)


# --------------------------------------------------------------------------------------- torch_geometric.nn
class DataParallel(torch.nn.Module):
    """``torch_geometric.nn.DataParallel`` @1.2.1 for ONE device per process: ``forward(data_list)`` batches the
    list and calls the module (train.py:280,285).  Multi-GPU runs use one process per GPU with
    ``cgc_net_b200.parallel`` (graphs sharded by node count exactly like PyG's ``scatter``) instead of threads."""

    def __init__(self, module, device_ids=None, output_device=None):
        super(DataParallel, self).__init__()
        self.module = module

    def forward(self, data_list):
        if len(data_list) == 0:
            raise ValueError('DataParallel received an empty data list')
        batch = Batch.from_data_list(data_list)
        p = next(self.module.parameters(), None)
        if p is not None and p.is_cuda:
            batch = batch.to(p.device, non_blocking=True)
            with torch.cuda.device(p.device):        # the library enqueues on the CURRENT device (see _backend._stream)
                return self.module(batch)
        return self.module(batch)


EXTRA_SELF_SLOT = False      # see radius_graph


def radius_graph(x, r, batch=None, loop=False, max_num_neighbors=32, flow='source_to_target'):
    """``torch_cluster.radius_graph`` @1.4.2.  CUDA positions run the library kernel (``graph.radius_graph``);
    CPU positions (loader workers) use scipy's cKDTree exactly like torch_cluster's CPU path.

    KNOWN UNCERTAINTY: with ``loop=True`` this returns at most
    ``max_num_neighbors`` in-edges per centre INCLUDING the self loop.  torch_cluster 1.4.x may instead query
    ``max_num_neighbors + 1`` candidates regardless of ``loop`` and only drop self pairs when ``loop=False``, which
    would give up to 9 in-edges (8 neighbours + self) for the reference's ``radius_graph(pos, 100, None, True, 8)``
    (dataflow/data.py:348).  The 1.4.2 source is neither vendored nor installable here, so it could not be
    checked; set ``EXTRA_SELF_SLOT = True`` to reproduce the k + 1 reading on both the CPU and the GPU path.  The
    model path is indifferent: it consumes whatever ``edge_index`` the loader produced."""
    if loop and EXTRA_SELF_SLOT:
        max_num_neighbors = int(max_num_neighbors) + 1
    assert flow in ('source_to_target', 'target_to_source')
    if x.is_cuda:
        ei = _graph.radius_graph(x, r, batch, loop, max_num_neighbors)
    else:
        import numpy as np
        from scipy.spatial import cKDTree
        counts = [x.size(0)] if batch is None else torch.bincount(batch).tolist()
        srcs, dsts, off = [], [], 0
        k = int(max_num_neighbors)
        for n in counts:
            if n == 0:
                continue
            p = x[off:off + n].double().numpy()
            kk = min(k + (0 if loop else 1), n)
            dist, idx = cKDTree(p).query(p, k=kk, distance_upper_bound=float(r))
            dist, idx = dist.reshape(n, kk), idx.reshape(n, kk)
            for c in range(n):
                cand = [(dist[c, s], int(idx[c, s])) for s in range(kk) if np.isfinite(dist[c, s])]
                if not loop:
                    cand = [t for t in cand if t[1] != c]
                for _, j in sorted(cand)[:k]:
                    srcs.append(off + j)
                    dsts.append(off + c)
            off += n
        ei = torch.tensor([srcs, dsts], dtype=torch.long)
    if flow == 'target_to_source':
        ei = ei.flip(0)
    return ei


# --------------------------------------------------------------------------------------- torch_geometric.utils
def sparse_to_dense(edge_index, edge_attr=None, num_nodes=None):
    n = int(edge_index.max()) + 1 if num_nodes is None else num_nodes
    adj = torch.zeros(n, n, dtype=torch.float, device=edge_index.device)
    adj[edge_index[0], edge_index[1]] = 1.0 if edge_attr is None else edge_attr
    return adj


def dense_to_sparse(adj):
    index = adj.nonzero().t().contiguous()
    return index, adj[index[0], index[1]]


def scatter_(name, src, index, dim_size=None):
    assert name == 'add', 'only scatter_("add") is used by the reference (model/utils.py:19)'
    if dim_size is None:
        dim_size = int(index.max()) + 1
    return src.new_zeros((dim_size,) + tuple(src.shape[1:])).index_add_(0, index, src)


def to_dense_batch(x, batch, fill_value=0):
    """1.2.x form: returns ``(dense_x, num_nodes)`` (the reference relies on it, network.py:242,258)."""
    num_nodes = scatter_('add', batch.new_ones(x.size(0)), batch)
    b, nmax = num_nodes.size(0), int(num_nodes.max())
    cum = torch.cat([batch.new_zeros(1), num_nodes.cumsum(0)])
    index = torch.arange(batch.size(0), device=x.device) - cum[batch] + batch * nmax
    out = x.new_full([b * nmax] + list(x.shape[1:]), fill_value)
    out[index] = x
    return out.view([b, nmax] + list(x.shape[1:])), num_nodes


# --------------------------------------------------------------------------------------- tensorboardX
class SummaryWriter(object):
    """No-op stand-in for ``tensorboardX.SummaryWriter`` (train.py:6); keeps the last scalar per tag."""

    def __init__(self, *args, **kwargs):
        self.scalars = {}

    def add_scalar(self, tag, value, step=None):
        self.scalars[tag] = (float(value), step)

    def close(self):
        pass


def install(replace_model=True, force=False):
    """Register the stand-ins (idempotent).  A real ``torch_geometric`` / ``tensorboardX`` already imported is
    left alone unless ``force``.  ``replace_model`` makes ``from model import network`` (train.py:16) resolve
    to ``cgc_net_b200.network``."""
    from .. import network

    def module(name, **attrs):
        m = types.ModuleType(name)
        m._cgc_compat = True
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    def wanted(name):
        cur = sys.modules.get(name)
        return force or cur is None or getattr(cur, '_cgc_compat', False) or getattr(cur, '_cgc_oracle_shim', False)

    if wanted('torch_geometric'):
        data = module('torch_geometric.data', Data=Data, Batch=Batch, Dataset=Dataset, DataLoader=DataLoader,
                      DataListLoader=DataListLoader)
        nn = module('torch_geometric.nn', DataParallel=DataParallel, radius_graph=radius_graph,
                    DenseSAGEConv=network.DenseSAGEConv, DenseGINConv=network.DenseGINConv)
        utils = module('torch_geometric.utils', sparse_to_dense=sparse_to_dense, dense_to_sparse=dense_to_sparse,
                       to_dense_batch=to_dense_batch, scatter_=scatter_)
        tg = module('torch_geometric', data=data, nn=nn, utils=utils, __version__='1.2.1+cgc_compat')
        sys.modules.update({'torch_geometric': tg, 'torch_geometric.data': data, 'torch_geometric.nn': nn,
                            'torch_geometric.utils': utils})
    if wanted('tensorboardX'):
        sys.modules['tensorboardX'] = module('tensorboardX', SummaryWriter=SummaryWriter)
    if replace_model:
        model = module('model', network=network)
        model.__path__ = []
        sys.modules['model'] = model
        sys.modules['model.network'] = network
