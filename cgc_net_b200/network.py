"""Drop-in ``SoftPoolingGcnEncoder`` for CGC-Net's ``model/network.py`` on B200.

Same public surface as the reference:

* constructor signature of ``SoftPoolingGcnEncoder`` (``model/network.py:128-131``; built at
  ``train.py:254-261``) and of ``GNN_Module`` (``network.py:58-59``);
* ``forward(data)`` -> ``(logits, loss)`` in training mode, ``logits`` in eval mode
  (``network.py:288-291``), for both input formats (PyG-style batch with ``x / edge_index /
  batch / y`` when ``load_data_sparse``; dense tuple ``(x, adj, num_nodes[, label])`` otherwise,
  ``network.py:251-256``);
* parameter / buffer names and shapes (``GCN_embed_1.gcn1.weight [in,out]``,
  ``...bn1.running_mean``, ``GCN_pool_1.lin.weight``, ``pred_model.0.weight`` ...), so reference
  checkpoints load with ``load_state_dict`` (``train.py:274``, ``common/utils.py:88-94``);
* ``assign_matrix`` (softmax ``S`` per pooling stage, detached) when ``collect_assign``
  (``network.py:141,201-202``, read at ``train.py:71``).

What differs is the execution: every graph operator runs as a hand-written sm_100a kernel
behind the C-ABI of ``include/cgc_sm100.h`` (see ``ops.py``); the stage-1 adjacency stays
sparse (CSR/CSC) instead of a dense ``[B, Nmax, Nmax]`` tensor.  The MLP head, cross-entropy,
dropout and the optimiser remain stock torch modules ("not behind the boundary").
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops

EPS = 1e-15


class DenseSAGEConv(nn.Module):
    """Parameter holder with PyG 1.2.1 ``DenseSAGEConv`` names, shapes and init
    (``weight [in, out]`` and ``bias [out]`` both ``U(-1/sqrt(in), 1/sqrt(in))``; one weight, no
    root weight).  The arithmetic runs inside ``ops.GnnModuleFunction``."""

    def __init__(self, in_channels, out_channels, normalize=True, bias=True):
        super(DenseSAGEConv, self).__init__()
        assert normalize, 'the reference always builds DenseSAGEConv(normalize=True) (network.py:95)'
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = nn.Parameter(torch.empty(in_channels, out_channels))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        bound = 1.0 / math.sqrt(self.in_channels)
        with torch.no_grad():
            self.weight.uniform_(-bound, bound)
            if self.bias is not None:
                self.bias.uniform_(-bound, bound)


class DenseGINConv(nn.Module):
    """Parameter holder with PyG 1.2.1 ``DenseGINConv`` names (``nn.0.*``, ``nn.2.*`` of the 2-layer MLP built at
    network.py:97-98, and the ``eps`` buffer).  With ``add_loop=False`` (every reference call) ``eps`` is unused:
    ``out = MLP(adj @ x) * mask``.  The arithmetic runs through ``ops`` (see ``GNN_Module._forward_gin``)."""

    def __init__(self, in_channels, out_channels, eps=0.0):
        super(DenseGINConv, self).__init__()
        # the middle module is the activation in the reference (no parameters): an Identity placeholder keeps
        # the Sequential indices 0 / 2 of the two Linear layers
        self.nn = nn.Sequential(nn.Linear(in_channels, out_channels), nn.Identity(),
                                nn.Linear(out_channels, out_channels))
        self.register_buffer('eps', torch.tensor([float(eps)]))


class GNN_Module(nn.Module):
    """Three conv -> activation -> BatchNorm1d blocks, concat, optional Linear (network.py:57-125); conv is
    DenseSAGEConv (``gcn_name='SAGE'``) or DenseGINConv around a 2-layer MLP (anything else, network.py:92-99)."""

    def __init__(self, input_dim, hidden_dim, embedding_dim, bias, bn, add_loop, lin=True,
                 gcn_name='SAGE', sync=False, activation='relu', jk=False):
        super(GNN_Module, self).__init__()
        if activation not in ops.ACTIVATIONS:
            raise ValueError('activation must be one of %s (network.py:83)' % sorted(ops.ACTIVATIONS))
        if add_loop:
            raise NotImplementedError('add_loop=True is never used by the reference '
                                      '(network.py:144,148,157,161,164)')
        self.jk = jk
        self.add_loop = add_loop
        self.act = ops.ACTIVATIONS[activation]
        self.gin = gcn_name != 'SAGE'
        if self.gin:
            self.gcn1 = DenseGINConv(input_dim, hidden_dim)
            self.gcn2 = DenseGINConv(hidden_dim, hidden_dim)
            self.gcn3 = DenseGINConv(hidden_dim, embedding_dim)
        else:
            self.gcn1 = DenseSAGEConv(input_dim, hidden_dim, normalize=True, bias=bias)
            self.gcn2 = DenseSAGEConv(hidden_dim, hidden_dim, normalize=True, bias=bias)
            self.gcn3 = DenseSAGEConv(hidden_dim, embedding_dim, normalize=True, bias=bias)
        self.has_bn = bool(bn)
        if bn:
            self.bn1 = nn.BatchNorm1d(hidden_dim)
            self.bn2 = nn.BatchNorm1d(hidden_dim)
            self.bn3 = nn.BatchNorm1d(embedding_dim)
        if lin is True:
            self.lin = nn.Linear(2 * hidden_dim + embedding_dim, embedding_dim)
        else:
            self.lin = None

    def forward(self, x, adj, mask=None, graph=None, num_graphs=None, nmax=None, scale=None, agg_cache=None):
        """``x [M, Cin]`` in the padded node layout.  Either ``graph`` (``ops.SparseGraph``,
        stage 1) or a dense ``adj [B, N, N]`` with ``scale = ops.dense_row_scale(adj)``; ``agg_cache`` = a dict shared
        by the modules of one dense stage (the first layer's ``A x`` is computed once)."""
        if self.gin:
            return self._forward_gin(x, adj, mask, graph, num_graphs, nmax)
        params, buffers = [], []
        for k in (1, 2, 3):
            conv = getattr(self, 'gcn%d' % k)
            params += [conv.weight, conv.bias]
            if self.has_bn:
                bn = getattr(self, 'bn%d' % k)
                params += [bn.weight, bn.bias]
                buffers.append((bn.running_mean, bn.running_var))
                if self.training:
                    bn.num_batches_tracked += 1
            else:
                params += [None, None]
        params += [self.lin.weight, self.lin.bias] if self.lin is not None else [None, None]
        if graph is not None:
            cfg = ops.GnnConfig(graph=graph, num_graphs=graph.num_graphs, nmax=graph.nmax,
                                mask=graph.mask, training=self.training, has_bn=self.has_bn,
                                has_lin=self.lin is not None, bn_buffers=buffers, act=self.act)
            return ops.GnnModuleFunction.apply(cfg, x, None, graph.rs, None, *params)
        rs, rsum = scale
        cfg = ops.GnnConfig(graph=None, num_graphs=num_graphs, nmax=nmax, mask=mask,
                            training=self.training, has_bn=self.has_bn,
                            has_lin=self.lin is not None, bn_buffers=buffers, act=self.act, agg_cache=agg_cache)
        return ops.GnnModuleFunction.apply(cfg, x, adj, rs, rsum, *params)

    def _forward_gin(self, x, adj, mask, graph, num_graphs, nmax):
        """DenseGINConv path (network.py:96-99,109-125): ``x_k = BN_k(act(MLP_k(A x_{k-1}) * mask))``.  Built
        from the same C-ABI kernels (neighbour-sum gather / contraction, tcgen05 or fp32 GEMM with fused bias
        and row mask, activation, BatchNorm), chained by autograd."""
        if graph is not None:
            num_graphs, nmax, mask = graph.num_graphs, graph.nmax, graph.mask
        cfg = ops.GnnConfig(graph=graph, num_graphs=num_graphs, nmax=nmax, mask=mask, training=self.training)
        feats = []
        h = x
        for k in (1, 2, 3):
            conv = getattr(self, 'gcn%d' % k)
            agg = ops.NeighbourSumFunction.apply(cfg, h, adj)
            z = ops.LinearFunction.apply(agg, conv.nn[0].weight, conv.nn[0].bias, None)
            z = ops.ActFunction.apply(z, self.act)
            z = ops.LinearFunction.apply(z, conv.nn[2].weight, conv.nn[2].bias, mask)
            z = ops.ActFunction.apply(z, self.act)
            if self.has_bn:
                bn = getattr(self, 'bn%d' % k)
                if self.training:
                    bn.num_batches_tracked += 1
                z = ops.BatchNormFunction.apply(z, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                                self.training, mask)
            feats.append(z)
            h = z
        out = torch.cat(feats, dim=1)
        if self.lin is not None:
            out = ops.LinearFunction.apply(out, self.lin.weight, self.lin.bias, mask)
        return out


class DenseJK(nn.Module):
    """Jumping-knowledge aggregation of the three layer outputs of one stage (reference ``DenseJK('lstm')``,
    network.py:11-55): a bi-LSTM over the 3-step "sequence" of per-layer features of every node scores the
    layers, the scores are soft-maxed and the layers summed with those weights.

    SURVEY 8(f) rank-1 row: for 3 layers of 16 / 20 / 32 channels (``--hidden-dim``, train.py:393; 20 is the
    reference's default) the whole aggregation is ONE hand-written kernel per direction of autograd
    (``ops.DenseJKFunction`` -> ``cgc_jk_lstm_fwd / bwd``); wider shapes, whose LSTM weights do not fit in shared
    memory (64 channels: 491 KB), keep torch's LSTM (cuDNN, TF32 off).  Parameter names (``lstm.*``, ``att.*``) match the reference for
    checkpoint compatibility (the ``nn.LSTM`` / ``nn.Linear`` modules are kept as parameter holders).
    """

    def __init__(self, mode, channels=None, num_layers=None):
        super(DenseJK, self).__init__()
        assert mode.lower() == 'lstm' and channels is not None and num_layers is not None
        self.mode, self.channel = mode.lower(), channels
        hidden = channels * num_layers // 2
        self.lstm = nn.LSTM(channels, hidden, bidirectional=True, batch_first=True)
        self.att = nn.Linear(2 * hidden, 1)

    def forward(self, feats, num_graphs, nmax):
        """``feats [B*N, L*C]`` (concat of the L layer outputs, padded layout) -> ``[B*N, C]``."""
        from . import _backend
        be = _backend.get()
        layers = feats.shape[1] // self.channel
        if be.jk_supported(self.channel, layers):
            # fused per-node kernel (cgc_jk_lstm_fwd / bwd): no cuDNN launch, no [rows, L, 2H] intermediates
            names = ['weight_ih_l0', 'weight_hh_l0', 'bias_ih_l0', 'bias_hh_l0']
            params = [getattr(self.lstm, n) for n in names] + [getattr(self.lstm, n + '_reverse') for n in names]
            return ops.DenseJKFunction.apply(feats, self.channel, layers, self.att.weight, self.att.bias, *params)
        rows = feats.shape[0]
        seq = feats.view(rows, -1, self.channel)                 # [rows, L, C]
        # fp32 parity with the reference: cuDNN's RNN would otherwise run its matmuls in single-pass TF32
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            score, _ = self.lstm(seq)
        alpha = torch.softmax(self.att(score).squeeze(-1), dim=-1)
        return (seq * alpha.unsqueeze(-1)).sum(dim=1)

    def __repr__(self):
        return '{}({})'.format(self.__class__.__name__, self.mode)


class SoftPoolingGcnEncoder(nn.Module):
    """CGC-Net's 3-stage DiffPool encoder (network.py:127-291) on hand-written sm_100a kernels."""

    def __init__(self, max_num_nodes, input_dim, hidden_dim, embedding_dim, bias, bn, assign_hidden_dim,
                 label_dim, assign_ratio=0.25, pred_hidden_dims=[50], concat=True, gcn_name='SAGE',
                 collect_assign=False, load_data_sparse=False, norm_adj=False,
                 activation='relu', drop_out=0., jk=False):
        super(SoftPoolingGcnEncoder, self).__init__()
        self.jk = jk
        self.drop_out = drop_out
        self.norm_adj = norm_adj
        self.load_data_sparse = load_data_sparse
        self.collect_assign = collect_assign
        self.assign_matrix = []
        self.assign_index = []          # argmax(S) per pooling stage (cluster id per node)
        assign_dim = int(max_num_nodes * assign_ratio)
        kw = dict(gcn_name=gcn_name, activation=activation, jk=jk)
        self.GCN_embed_1 = GNN_Module(input_dim, hidden_dim, embedding_dim, bias, bn,
                                      add_loop=False, lin=False, **kw)
        if jk:
            self.jk1 = DenseJK('lstm', hidden_dim, 3)
        self.GCN_pool_1 = GNN_Module(input_dim, assign_hidden_dim, assign_dim, bias, bn,
                                     add_loop=False, **kw)
        if concat and not jk:
            input_dim = hidden_dim * 2 + embedding_dim
        else:
            input_dim = embedding_dim
        assign_dim = int(assign_dim * assign_ratio)
        self.GCN_embed_2 = GNN_Module(input_dim, hidden_dim, embedding_dim, bias, bn,
                                      add_loop=False, lin=False, **kw)
        if jk:
            self.jk2 = DenseJK('lstm', hidden_dim, 3)
        self.GCN_pool_2 = GNN_Module(input_dim, assign_hidden_dim, assign_dim, bias, bn,
                                     add_loop=False, **kw)
        self.GCN_embed_3 = GNN_Module(input_dim, hidden_dim, embedding_dim, bias, bn,
                                      add_loop=False, lin=False, **kw)
        if jk:
            self.jk3 = DenseJK('lstm', hidden_dim, 3)
        pred_input = input_dim * 3
        self.pred_model = self.build_readout_module(pred_input, pred_hidden_dims, label_dim, activation)

    def build_readout_module(self, pred_input_dim, pred_hidden_dims, label_dim, activation):
        """MLP head; stays a stock torch module (network.py:220-234)."""
        if len(pred_hidden_dims) == 0:
            return nn.Linear(pred_input_dim, label_dim)
        layers = []
        for pred_dim in pred_hidden_dims:
            layers.append(nn.Linear(pred_input_dim, pred_dim))
            layers.append({'relu': nn.ReLU, 'elu': nn.ELU, 'leakyrelu': nn.LeakyReLU}[activation](inplace=True))
            pred_input_dim = pred_dim
            if self.drop_out > 0:
                layers.append(nn.Dropout(self.drop_out))
        layers.append(nn.Linear(pred_dim, label_dim))
        return nn.Sequential(*layers)

    # ------------------------------------------------------------------ input handling
    def _sparse_input(self, data):
        """network.py:237-243 without the dense adjacency."""
        counts = getattr(data, 'num_nodes_list', None)
        xp, graph = ops.build_sparse_graph(data.x, data.edge_index, data.batch, counts, self.norm_adj,
                                           max_degree=getattr(data, 'max_degree', None),
                                           self_loops=getattr(data, 'self_loops', None))
        return xp, graph, data.y

    # ------------------------------------------------------------------ forward
    def forward(self, data):
        self.assign_matrix = []
        self.assign_index = []
        readouts = []
        label = None
        if self.load_data_sparse:
            x, graph, label = self._sparse_input(data)
            b, n, mask = graph.num_graphs, graph.nmax, graph.mask
            adj = scale = None
        else:
            x3, adj, counts = data[0], data[1], data[2]
            if self.training:
                label = data[3]
            b, n = adj.shape[0], adj.shape[1]
            graph = None
            num_nodes = torch.as_tensor(counts).to(device=x3.device, dtype=torch.int32)
            mask = ops.RowMask(num_nodes, n)
            x = x3.reshape(b * n, -1).contiguous().float()
            adj = adj.contiguous().float()
            if self.norm_adj:
                adj = ops.DenseRenormFunction.apply(adj, ops.RENORM_P, mask)
            scale = ops.dense_row_scale(adj)

        sage = not self.GCN_embed_1.gin
        for stage in (1, 2, 3):
            embed = getattr(self, 'GCN_embed_%d' % stage)
            kw = dict(agg_cache={}) if (sage and graph is None) else {}      # dense stage: A x of layer 1 shared
            feat = embed(x, adj, mask, graph=graph, num_graphs=b, nmax=n, scale=scale, **kw)
            if self.jk:
                feat = getattr(self, 'jk%d' % stage)(feat, b, n)
            pooled, _ = ops.ColMaxFunction.apply(feat, b, n)
            readouts.append(pooled)
            if stage == 3:
                break
            pool = getattr(self, 'GCN_pool_%d' % stage)
            assign = pool(x, adj, mask, graph=graph, num_graphs=b, nmax=n, scale=scale, **kw)
            cfg = ops.GnnConfig(graph=graph, num_graphs=b, nmax=n, mask=mask)
            xp, ap, s_full, arg = ops.DiffPoolFunction.apply(cfg, assign, feat, adj, self.collect_assign)
            if self.collect_assign:
                self.assign_matrix.append(s_full.view(b, n, -1))
                self.assign_index.append(arg.view(b, n))
            n = xp.shape[1]
            x = xp.reshape(b * n, -1)
            adj, graph, mask = ap, None, None
            if self.norm_adj:
                adj = ops.DenseRenormFunction.apply(adj, ops.RENORM_P, None)
            scale = ops.dense_row_scale(adj)

        output = self.pred_model(torch.cat(readouts, 1))
        if self.training:
            return output, F.cross_entropy(output, label)
        return output
