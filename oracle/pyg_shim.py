"""Minimal stand-in for the ``torch_geometric==1.2.1`` symbols CGC-Net imports.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  ``torch_geometric`` / ``torch_scatter``
are pinned in ``/root/reference/requirements.txt:40-41`` but are neither vendored under
``/root/reference`` nor installable offline, so the four symbols used on the hot path are
restated here from the published 1.2.1 behaviour:

* ``torch_geometric.nn.DenseSAGEConv``   (call sites ``model/network.py:3,95,114-116``)
* ``torch_geometric.nn.DenseGINConv``    (``model/network.py:3,99``)
* ``torch_geometric.utils.to_dense_batch`` (``model/network.py:5,242``) -- 1.2.x returns
  ``(dense_x, num_nodes)``, which the reference relies on at ``network.py:242,258``
* ``torch_geometric.utils.scatter_``     (``model/utils.py:2,19``)

``install()`` registers them in ``sys.modules`` so the reference files import unmodified.
"""
import math
import sys
import types

import torch
import torch.nn.functional as F
from torch.nn import Parameter


def _uniform(size, tensor):
    # torch_geometric/nn/inits.py::uniform @1.2.1
    bound = 1.0 / math.sqrt(size)
    if tensor is not None:
        tensor.data.uniform_(-bound, bound)


class DenseSAGEConv(torch.nn.Module):
    """``torch_geometric/nn/dense/dense_sage_conv.py`` @1.2.1: single weight, no root weight."""

    def __init__(self, in_channels, out_channels, normalize=True, bias=True):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.normalize = normalize
        self.weight = Parameter(torch.Tensor(self.in_channels, out_channels))
        if bias:
            self.bias = Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        _uniform(self.in_channels, self.weight)
        _uniform(self.in_channels, self.bias)

    def forward(self, x, adj, mask=None, add_loop=True):
        x = x.unsqueeze(0) if x.dim() == 2 else x
        adj = adj.unsqueeze(0) if adj.dim() == 2 else adj
        B, N, _ = x.size()
        if add_loop:
            adj = adj.clone()
            idx = torch.arange(N, dtype=torch.long, device=adj.device)
            adj[:, idx, idx] = 1
        out = torch.matmul(adj, x)
        out = out / adj.sum(dim=-1, keepdim=True).clamp(min=1)
        out = torch.matmul(out, self.weight)
        if self.bias is not None:
            out = out + self.bias
        if self.normalize:
            out = F.normalize(out, p=2, dim=-1)
        if mask is not None:
            out = out * mask.view(B, N, 1).to(x.dtype)
        return out


class DenseGINConv(torch.nn.Module):
    """``torch_geometric/nn/dense/dense_gin_conv.py`` @1.2.1."""

    def __init__(self, nn, eps=0, train_eps=False):
        super().__init__()
        self.nn = nn
        self.initial_eps = eps
        if train_eps:
            self.eps = Parameter(torch.Tensor([eps]))
        else:
            self.register_buffer('eps', torch.Tensor([eps]))
        self.reset_parameters()

    def reset_parameters(self):
        for m in self.nn.modules():
            if hasattr(m, 'reset_parameters') and m is not self.nn:
                m.reset_parameters()
        self.eps.data.fill_(self.initial_eps)

    def forward(self, x, adj, mask=None, add_loop=True):
        x = x.unsqueeze(0) if x.dim() == 2 else x
        adj = adj.unsqueeze(0) if adj.dim() == 2 else adj
        B, N, _ = x.size()
        out = torch.matmul(adj, x)
        if add_loop:
            out = (1 + self.eps) * x + out
        out = self.nn(out)
        if mask is not None:
            out = out * mask.view(B, N, 1).to(x.dtype)
        return out


def scatter_(name, src, index, dim_size=None):
    """``torch_geometric/utils/scatter.py`` @1.2.1 -> ``torch_scatter.scatter_add`` (dim 0)."""
    assert name in ('add', 'mean', 'max')
    if name != 'add':
        raise NotImplementedError('only scatter_("add") is on the reference hot path')
    if dim_size is None:
        dim_size = int(index.max().item()) + 1
    out = src.new_zeros((dim_size,) + tuple(src.shape[1:]))
    return out.index_add_(0, index, src)


def to_dense_batch(x, batch, fill_value=0):
    """``torch_geometric/utils/to_dense_batch.py`` @1.2.x: returns ``(dense_x, num_nodes)``."""
    num_nodes = scatter_('add', batch.new_ones(x.size(0)), batch)
    batch_size, max_num_nodes = num_nodes.size(0), int(num_nodes.max().item())
    cum_nodes = torch.cat([batch.new_zeros(1), num_nodes.cumsum(dim=0)], dim=0)
    index = torch.arange(batch.size(0), dtype=torch.long, device=x.device)
    index = (index - cum_nodes[batch]) + (batch * max_num_nodes)
    size = [batch_size * max_num_nodes] + list(x.size())[1:]
    batch_x = x.new_full(size, fill_value)
    batch_x[index] = x
    size = [batch_size, max_num_nodes] + list(x.size())[1:]
    batch_x = batch_x.view(size)
    return batch_x, num_nodes


def install():
    """Register the shim under the ``torch_geometric`` module names (idempotent)."""
    if 'torch_geometric' in sys.modules and not getattr(
            sys.modules['torch_geometric'], '_cgc_oracle_shim', False):
        return  # a real torch_geometric is importable: leave it alone
    tg = types.ModuleType('torch_geometric')
    tg._cgc_oracle_shim = True
    tg_nn = types.ModuleType('torch_geometric.nn')
    tg_utils = types.ModuleType('torch_geometric.utils')
    tg_nn.DenseSAGEConv = DenseSAGEConv
    tg_nn.DenseGINConv = DenseGINConv
    tg_utils.to_dense_batch = to_dense_batch
    tg_utils.scatter_ = scatter_
    tg.nn = tg_nn
    tg.utils = tg_utils
    sys.modules['torch_geometric'] = tg
    sys.modules['torch_geometric.nn'] = tg_nn
    sys.modules['torch_geometric.utils'] = tg_utils
