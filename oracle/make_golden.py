"""Mint ``tests/golden/*.npz`` from the VERBATIM reference.  TEST INFRASTRUCTURE.

Run in the build container only (needs the read-only checkout at /root/reference):

    python -m oracle.make_golden

The reference ships no tests, fixtures or golden vectors, so the known
answers are produced by importing ``/root/reference/model/network.py`` and
``model/utils.py`` unmodified behind the PyG shim (``oracle/reference_loader.py``) and
recording inputs, parameters and outputs for fixed seeds:

* ``sage_conv_n200_k5``       BASELINE config #1: one 200-node k=5 cell graph, 16-d features,
                              one DenseSAGEConv(16 -> 20) on ``to_dense_adj`` output (CPU).
* ``pool_stage1_small``       BASELINE config #2 (reduced): GCN_embed_1 + GCN_pool_1 + _diff_pool.
* ``encoder_small_plain`` /   BASELINE config #3 (reduced): full SoftPoolingGcnEncoder forward +
  ``encoder_small_normadj``   backward on a ragged 4-graph batch with duplicate edges and an
                              isolated row; logits, loss, every parameter gradient, BatchNorm
                              running statistics, softmax(S) arg-max per pooling stage.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cgc_net_b200 import graph  # noqa: E402  (host-side synthetic graph generator only)
from oracle import reference_loader  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


def _np(t):
    return t.detach().cpu().numpy()


def _messy_graphs(sizes, degrees, feat, seed):
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import helpers
    return helpers.make_graphs(sizes, degrees, feat, seed, messy=True)


def golden_sage_conv():
    ref = reference_loader.load_network_module()
    from model.utils import to_dense_adj            # verbatim reference (model/utils.py)
    d = graph.synthetic_cell_graph(200, 5, 16, seed=7, row_is_centre=False)
    batch = torch.zeros(200, dtype=torch.long)
    adj = to_dense_adj(d.edge_index, batch)
    torch.manual_seed(3)
    conv = ref.DenseSAGEConv(16, 20, normalize=True, bias=True)
    out = conv(d.x.unsqueeze(0), adj, None, False)
    np.savez_compressed(os.path.join(OUT, 'sage_conv_n200_k5.npz'), x=_np(d.x),
                        edge_index=_np(d.edge_index), weight=_np(conv.weight), bias=_np(conv.bias),
                        adj_rowsum=_np(adj.sum(-1)[0]), out=_np(out[0]))


def _encoder_case(name, norm_adj, sizes, degrees, max_nodes, ratio, seed, gcn_name='SAGE', activation='relu', jk=False):
    graphs = _messy_graphs(sizes, degrees, 16, seed)
    batch = graph.Batch.from_data_list(graphs)
    args = (max_nodes, 16, 20, 20, True, True, 20, 3, ratio, [50])
    kw = dict(concat=True, gcn_name=gcn_name, collect_assign=True, load_data_sparse=True,
              norm_adj=norm_adj, activation=activation, jk=jk)
    torch.manual_seed(seed)
    model = reference_loader.build_reference_encoder(*args, **kw)
    model.train()
    state = {k: _np(v).copy() for k, v in model.state_dict().items()}
    logits, loss = reference_loader.run_reference(model, batch)
    loss.backward()
    rec = dict(x=_np(batch.x), edge_index=_np(batch.edge_index), batch=_np(batch.batch), y=_np(batch.y),
               num_nodes=np.asarray(batch.num_nodes_list), max_nodes=max_nodes, ratio=ratio,
               norm_adj=int(norm_adj), logits=_np(logits), loss=_np(loss), gcn_name=np.asarray(gcn_name),
               activation=np.asarray(activation), jk=int(jk))
    for k, v in state.items():
        rec['param/' + k] = v
    for k, p in model.named_parameters():
        rec['grad/' + k] = _np(p.grad)
    for k, v in model.state_dict().items():
        if 'running_' in k:
            rec['after/' + k] = _np(v)
    for i, s in enumerate(model.assign_matrix):
        rec['assign_argmax/%d' % i] = _np(s.argmax(-1)).astype(np.int32)
        rec['assign_top2_margin/%d' % i] = _np(s.topk(2, dim=-1).values.diff(dim=-1).abs().squeeze(-1))
        rec['assign_colsum/%d' % i] = _np(s.sum(1))
    model.eval()
    with torch.no_grad():
        rec['eval_logits'] = _np(reference_loader.run_reference(model, batch))
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **rec)
    print(name, 'logits', rec['logits'].ravel()[:3], 'loss', float(rec['loss']))


def golden_pool_stage1():
    ref = reference_loader.load_network_module()
    from model.utils import to_dense_adj
    graphs = _messy_graphs([90, 120, 75], [8, 8, 6], 16, seed=5)
    batch = graph.Batch.from_data_list(graphs)
    torch.manual_seed(11)
    model = reference_loader.build_reference_encoder(300, 16, 20, 20, True, True, 20, 3, 0.1, [50],
                                                     concat=True, load_data_sparse=True)
    model.train()
    state = {k: _np(v).copy() for k, v in model.state_dict().items()}
    x, adj, counts, _ = model._sparse_to_dense_input(batch)
    mask = model.construct_mask(adj.shape[1], counts)
    emb = model.GCN_embed_1(x, adj, mask)
    assign = model.GCN_pool_1(x, adj, mask)
    xp, ap = model._diff_pool(emb, adj, assign, mask)
    rec = dict(x=_np(batch.x), edge_index=_np(batch.edge_index), batch=_np(batch.batch),
               num_nodes=np.asarray(batch.num_nodes_list), embed=_np(emb), assign=_np(assign),
               pooled_x=_np(xp), pooled_adj=_np(ap))
    for k, v in state.items():
        rec['param/' + k] = v
    np.savez_compressed(os.path.join(OUT, 'pool_stage1_small.npz'), **rec)


def main():
    os.makedirs(OUT, exist_ok=True)
    assert reference_loader.available(), 'needs /root/reference'
    golden_sage_conv()
    golden_pool_stage1()
    _encoder_case('encoder_small_plain', False, [150, 97, 181, 64], [8, 5, 8, 6], 500, 0.1, seed=21)
    _encoder_case('encoder_small_normadj', True, [150, 97, 181, 64], [8, 5, 8, 6], 500, 0.1, seed=22)
    # variants of the reference (network.py:82-99, 11-55): DenseGINConv + elu, and the published --jk configuration
    _encoder_case('encoder_small_gin_elu', True, [120, 97, 64], [8, 5, 6], 400, 0.1, seed=23, gcn_name='GIN',
                  activation='elu')
    _encoder_case('encoder_small_jk', False, [120, 97, 64], [8, 5, 6], 400, 0.1, seed=24, jk=True)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == '__main__':
    main()
