"""Dense pure-torch CPU restatement of the CGC-Net hot path.  TEST INFRASTRUCTURE.

Self-contained (no torch_geometric, no /root/reference): this is the oracle that travels
to the GPU box and the thing ``bench.py --impl reference`` times on the host cores.  It
performs the same dense ``[B, Nmax, Nmax]`` arithmetic, in the same order, as

* ``model/network.py:57-125``   GNN_Module            -> ``OracleGNNModule``
* ``model/network.py:127-291``  SoftPoolingGcnEncoder -> ``OracleEncoder``
* ``model/network.py:11-55``    DenseJK               -> ``OracleJK``
* ``model/utils.py:3-37``       to_dense_adj          -> ``densify_edges``
* torch_geometric 1.2.1 ``DenseSAGEConv`` / ``DenseGINConv`` / ``to_dense_batch``
  (un-vendored; restated from the pinned version, see ``oracle/pyg_shim.py``)

Parameter and buffer names equal the reference's, so ``state_dict``s are interchangeable
between the verbatim reference, this oracle and the CUDA module.  It is checked against
the verbatim import (``tests/test_oracle.py``) and pinned by ``tests/golden/*.npz``.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

RENORM_EPS = 1e-15   # model/network.py:8
RENORM_P = 0.4       # model/network.py:260,271,280


# --------------------------------------------------------------------------- input side
def densify_edges(edge_index, batch):
    """COO -> dense 0/1 adjacency, ``adj[b, src_local, dst_local] = 1`` (model/utils.py:3-37).

    Duplicates collapse (assignment, not accumulation, utils.py:33); ``batch`` is sorted.
    """
    num_graphs = int(batch[-1]) + 1
    counts = torch.zeros(num_graphs, dtype=torch.long).index_add_(
        0, batch, torch.ones_like(batch))
    start = torch.cat([counts.new_zeros(1), counts.cumsum(0)])
    nmax = int(counts.max())
    adj = torch.zeros(num_graphs, nmax, nmax, dtype=torch.get_default_dtype())
    src, dst = edge_index[0], edge_index[1]
    g = batch[src]
    adj[g, src - start[g], dst - start[batch[dst]]] = 1.0
    return adj, counts


def pad_nodes(x, batch, counts):
    """``to_dense_batch`` @PyG 1.2.x: ``[sum N, F] -> [B, Nmax, F]`` zero padded."""
    num_graphs, nmax = counts.numel(), int(counts.max())
    start = torch.cat([counts.new_zeros(1), counts.cumsum(0)])
    slot = torch.arange(batch.numel()) - start[batch] + batch * nmax
    out = x.new_zeros((num_graphs * nmax,) + tuple(x.shape[1:]))
    out[slot] = x
    return out.view(num_graphs, nmax, *x.shape[1:])


def node_mask(nmax, counts):
    """``construct_mask`` (model/network.py:172-180) without the host loop: ``[B, Nmax, 1]``."""
    return (torch.arange(nmax)[None, :] < counts.view(-1, 1)).to(torch.get_default_dtype()).unsqueeze(2)


def renorm_adjacency(adj, p=RENORM_P, mask=None):
    """``_re_norm_adj`` (model/network.py:183-191).  NOTE: zeroes the diagonal of ``adj`` IN PLACE."""
    n = adj.shape[1]
    eye = torch.arange(n)
    adj[:, eye, eye] = 0
    out = adj / (adj.sum(-1, keepdim=True) + RENORM_EPS) * (1 - p)
    out[:, eye, eye] = p
    if mask is not None:
        out = out * mask
    return out


# --------------------------------------------------------------------------- conv layers
class OracleSAGEConv(nn.Module):
    """DenseSAGEConv @PyG 1.2.1 (single ``[in, out]`` weight, no root weight)."""

    def __init__(self, cin, cout, bias=True):
        super().__init__()
        bound = 1.0 / math.sqrt(cin)
        self.weight = nn.Parameter(torch.empty(cin, cout).uniform_(-bound, bound))
        if bias:
            self.bias = nn.Parameter(torch.empty(cout).uniform_(-bound, bound))
        else:
            self.register_parameter('bias', None)

    def forward(self, x, adj, mask, add_loop=False):
        if add_loop:
            adj = adj.clone()
            eye = torch.arange(adj.shape[1])
            adj[:, eye, eye] = 1
        h = adj @ x
        h = h / adj.sum(-1, keepdim=True).clamp(min=1)
        h = h @ self.weight
        if self.bias is not None:
            h = h + self.bias
        h = F.normalize(h, p=2, dim=-1)
        if mask is not None:
            h = h * mask
        return h


class OracleGINConv(nn.Module):
    """DenseGINConv @PyG 1.2.1 around the 2-layer MLP built at model/network.py:97-99."""

    def __init__(self, mlp):
        super().__init__()
        self.nn = mlp
        self.register_buffer('eps', torch.zeros(1))

    def forward(self, x, adj, mask, add_loop=False):
        h = adj @ x
        if add_loop:
            h = (1 + self.eps) * x + h
        h = self.nn(h)
        if mask is not None:
            h = h * mask
        return h


def _act(name):
    return {'relu': nn.ReLU, 'elu': nn.ELU, 'leakyrelu': nn.LeakyReLU}[name]()


class OracleGNNModule(nn.Module):
    """Three conv -> act -> BatchNorm1d blocks, concat, optional Linear (network.py:57-125)."""

    def __init__(self, input_dim, hidden_dim, embedding_dim, bias, bn, add_loop, lin=True,
                 gcn_name='SAGE', activation='relu'):
        super().__init__()
        self.add_loop = add_loop
        dims = [(input_dim, hidden_dim), (hidden_dim, hidden_dim), (hidden_dim, embedding_dim)]
        for k, (ci, co) in enumerate(dims, 1):
            if gcn_name == 'SAGE':
                conv = OracleSAGEConv(ci, co, bias)
            else:
                conv = OracleGINConv(nn.Sequential(nn.Linear(ci, co), _act(activation),
                                                   nn.Linear(co, co)))
            setattr(self, 'gcn%d' % k, conv)
            if bn:
                setattr(self, 'bn%d' % k, nn.BatchNorm1d(co))
        self.act = _act(activation)
        self.has_bn = bn
        self.lin = nn.Linear(2 * hidden_dim + embedding_dim, embedding_dim) if lin else None

    def _bn(self, k, x):
        if not self.has_bn:
            return x
        b, n, c = x.shape
        return getattr(self, 'bn%d' % k)(x.reshape(-1, c)).view(b, n, c)   # padded rows count

    def forward(self, x, adj, mask=None):
        feats = []
        for k in (1, 2, 3):
            x = self._bn(k, self.act(getattr(self, 'gcn%d' % k)(x, adj, mask, self.add_loop)))
            feats.append(x)
        out = torch.cat(feats, dim=-1)
        if mask is not None:
            out = out * mask
        if self.lin is not None:
            out = self.lin(out)
            if mask is not None:
                out = out * mask
        return out


class OracleJK(nn.Module):
    """DenseJK('lstm') (network.py:11-55): bi-LSTM attention over the three layer outputs."""

    def __init__(self, channels, num_layers):
        super().__init__()
        self.channel = channels
        self.lstm = nn.LSTM(channels, channels * num_layers // 2, bidirectional=True,
                            batch_first=True)
        self.att = nn.Linear(2 * channels * num_layers // 2, 1)

    def forward(self, xs):
        b, n, _ = xs.shape
        seq = torch.stack(torch.split(xs, self.channel, -1), 2).reshape(b * n, -1, self.channel)
        alpha, _ = self.lstm(seq)
        alpha = torch.softmax(self.att(alpha).squeeze(-1), dim=-1)
        return (seq * alpha.unsqueeze(-1)).sum(1).reshape(b, n, self.channel)


# --------------------------------------------------------------------------- encoder
def diff_pool(x, adj, assign, mask):
    """``_diff_pool`` (network.py:194-208). Returns (X', A', softmax(assign) before masking)."""
    s_full = torch.softmax(assign, dim=-1)
    s = s_full if mask is None else s_full * mask
    st = s.transpose(1, 2)
    return st @ x, (st @ adj) @ s, s_full


class OracleEncoder(nn.Module):
    """SoftPoolingGcnEncoder (network.py:127-291), same constructor signature."""

    def __init__(self, max_num_nodes, input_dim, hidden_dim, embedding_dim, bias, bn,
                 assign_hidden_dim, label_dim, assign_ratio=0.25, pred_hidden_dims=[50],
                 concat=True, gcn_name='SAGE', collect_assign=False, load_data_sparse=False,
                 norm_adj=False, activation='relu', drop_out=0., jk=False):
        super().__init__()
        self.jk, self.norm_adj = jk, norm_adj
        self.load_data_sparse, self.collect_assign = load_data_sparse, collect_assign
        self.assign_matrix = []
        kw = dict(gcn_name=gcn_name, activation=activation)
        c1 = int(max_num_nodes * assign_ratio)
        c2 = int(c1 * assign_ratio)
        self.GCN_embed_1 = OracleGNNModule(input_dim, hidden_dim, embedding_dim, bias, bn,
                                           False, lin=False, **kw)
        if jk:
            self.jk1 = OracleJK(hidden_dim, 3)
        self.GCN_pool_1 = OracleGNNModule(input_dim, assign_hidden_dim, c1, bias, bn, False, **kw)
        mid = hidden_dim * 2 + embedding_dim if (concat and not jk) else embedding_dim
        self.GCN_embed_2 = OracleGNNModule(mid, hidden_dim, embedding_dim, bias, bn, False,
                                           lin=False, **kw)
        if jk:
            self.jk2 = OracleJK(hidden_dim, 3)
        self.GCN_pool_2 = OracleGNNModule(mid, assign_hidden_dim, c2, bias, bn, False, **kw)
        self.GCN_embed_3 = OracleGNNModule(mid, hidden_dim, embedding_dim, bias, bn, False,
                                           lin=False, **kw)
        if jk:
            self.jk3 = OracleJK(hidden_dim, 3)
        layers, d = [], mid * 3
        for h in pred_hidden_dims:
            layers += [nn.Linear(d, h), _act(activation)]
            if drop_out > 0:
                layers.append(nn.Dropout(drop_out))
            d = h
        layers.append(nn.Linear(d, label_dim))
        self.pred_model = nn.Sequential(*layers) if pred_hidden_dims else layers[0]

    def _stage(self, k, x, adj, mask):
        emb = getattr(self, 'GCN_embed_%d' % k)(x, adj, mask)
        if self.jk:
            emb = getattr(self, 'jk%d' % k)(emb)
        return emb, emb.max(dim=1)[0]            # max over ALL Nmax rows (padded zeros included)

    def forward(self, data):
        self.assign_matrix = []
        if self.load_data_sparse:
            adj, counts = densify_edges(data.edge_index, data.batch)
            x = pad_nodes(data.x, data.batch, counts)
            label = data.y
        else:
            x, adj, counts = data[0], data[1], data[2]
            label = data[3] if self.training else None
        mask = node_mask(adj.shape[1], torch.as_tensor(counts).long())
        readouts = []
        for k in (1, 2, 3):
            if self.norm_adj:
                adj = renorm_adjacency(adj, RENORM_P, mask)
            emb, pooled = self._stage(k, x, adj, mask)
            readouts.append(pooled)
            if k < 3:
                assign = getattr(self, 'GCN_pool_%d' % k)(x, adj, mask)
                x, adj, s_full = diff_pool(emb, adj, assign, mask)
                if self.collect_assign:
                    self.assign_matrix.append(s_full.detach())
                mask = None
        logits = self.pred_model(torch.cat(readouts, 1))
        if self.training:
            return logits, F.cross_entropy(logits, label)
        return logits
