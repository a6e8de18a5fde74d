"""The oracle against (a) the committed golden vectors minted from the verbatim reference and
(b) the verbatim reference itself where /root/reference exists.  CPU only."""
import os

import numpy as np
import pytest
import torch

import helpers
from cgc_net_b200 import graph
from oracle import pyg_shim, reference_loader, restatement


def _load(name):
    return np.load(os.path.join(helpers.GOLDEN_DIR, name + '.npz'))


def _batch_from(g):
    return graph.Batch(torch.from_numpy(g['x']), torch.from_numpy(g['edge_index']),
                       torch.from_numpy(g['batch']), torch.from_numpy(g['y']) if 'y' in g else None,
                       [int(n) for n in g['num_nodes']])


GOLDEN_ENCODERS = ['encoder_small_plain', 'encoder_small_normadj', 'encoder_small_gin_elu', 'encoder_small_jk']


def _variant_kwargs(g):
    """Constructor keywords recorded with a golden encoder case (oracle/make_golden.py)."""
    return dict(gcn_name=str(g['gcn_name']) if 'gcn_name' in g.files else 'SAGE',
                activation=str(g['activation']) if 'activation' in g.files else 'relu',
                jk=bool(int(g['jk'])) if 'jk' in g.files else False, norm_adj=bool(int(g['norm_adj'])))


def _oracle_from(g, **kw):
    args = (int(g['max_nodes']), 16, 20, 20, True, True, 20, 3, float(g['ratio']), [50])
    m = restatement.OracleEncoder(*args, concat=True, load_data_sparse=True, **_variant_kwargs(g), **kw)
    m.load_state_dict({k[len('param/'):]: torch.from_numpy(g[k]) for k in g.files if k.startswith('param/')})
    return m


def test_sage_conv_golden():
    """BASELINE config #1: single 200-node k=5 graph, one DenseSAGEConv(16->20)."""
    g = _load('sage_conv_n200_k5')
    adj, counts = restatement.densify_edges(torch.from_numpy(g['edge_index']), torch.zeros(200, dtype=torch.long))
    assert np.array_equal(adj.sum(-1)[0].numpy(), g['adj_rowsum'])          # bit-exact index work
    conv = restatement.OracleSAGEConv(16, 20)
    conv.weight.data = torch.from_numpy(g['weight'])
    conv.bias.data = torch.from_numpy(g['bias'])
    out = conv(torch.from_numpy(g['x']).unsqueeze(0), adj, None)
    np.testing.assert_allclose(out[0].detach().numpy(), g['out'], rtol=0, atol=1e-6)


@pytest.mark.parametrize('name', GOLDEN_ENCODERS)
def test_encoder_golden(name):
    g = _load(name)
    m = _oracle_from(g, collect_assign=True)
    m.train()
    logits, loss = m(_batch_from(g))
    np.testing.assert_allclose(logits.detach().numpy(), g['logits'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(float(loss), float(g['loss']), rtol=1e-6)
    loss.backward()
    for k, p in m.named_parameters():
        ref = g['grad/' + k]
        assert np.abs(p.grad.numpy() - ref).max() <= 1e-5 * max(np.abs(ref).max(), 1e-3), k
    for k, v in m.state_dict().items():
        if 'running_' in k:
            np.testing.assert_allclose(v.numpy(), g['after/' + k], rtol=1e-5, atol=1e-7)
    for i, s in enumerate(m.assign_matrix):
        mine = s.argmax(-1).numpy()
        safe = g['assign_top2_margin/%d' % i] > 1e-6
        assert np.array_equal(mine[safe], g['assign_argmax/%d' % i][safe])
    m.eval()
    with torch.no_grad():
        np.testing.assert_allclose(m(_batch_from(g)).numpy(), g['eval_logits'], rtol=1e-5, atol=1e-6)


def test_pool_stage1_golden():
    """BASELINE config #2 (reduced): GCN_embed_1 + GCN_pool_1 + _diff_pool."""
    g = _load('pool_stage1_small')
    m = restatement.OracleEncoder(300, 16, 20, 20, True, True, 20, 3, 0.1, [50], load_data_sparse=True)
    m.load_state_dict({k[len('param/'):]: torch.from_numpy(g[k]) for k in g.files if k.startswith('param/')})
    m.train()
    b = _batch_from(g)
    adj, counts = restatement.densify_edges(b.edge_index, b.batch)
    x = restatement.pad_nodes(b.x, b.batch, counts)
    mask = restatement.node_mask(adj.shape[1], counts)
    emb = m.GCN_embed_1(x, adj, mask)
    assign = m.GCN_pool_1(x, adj, mask)
    xp, ap, _ = restatement.diff_pool(emb, adj, assign, mask)
    for mine, key in ((emb, 'embed'), (assign, 'assign'), (xp, 'pooled_x'), (ap, 'pooled_adj')):
        np.testing.assert_allclose(mine.detach().numpy(), g[key], rtol=1e-5, atol=1e-6)


@pytest.mark.skipif(not reference_loader.available(), reason='verbatim reference not on this host')
@pytest.mark.parametrize('norm_adj', [False, True])
def test_oracle_equals_verbatim_reference(norm_adj):
    graphs = helpers.make_graphs([70, 41, 55], [8, 5, 6], seed=3)
    batch = graph.Batch.from_data_list(graphs)
    args = helpers.encoder_args(300)
    kw = dict(concat=True, gcn_name='SAGE', load_data_sparse=True, norm_adj=norm_adj)
    torch.manual_seed(5)
    ref = reference_loader.build_reference_encoder(*args, **kw)
    orc = restatement.OracleEncoder(*args, **kw)
    orc.load_state_dict(ref.state_dict())
    ref.train(), orc.train()
    lr, lossr = reference_loader.run_reference(ref, batch)
    lo, losso = orc(batch)
    assert torch.equal(lr, lo) or helpers.rel_err(lo, lr) < 1e-6
    lossr.backward(), losso.backward()
    rg = dict(ref.named_parameters())
    for k, p in orc.named_parameters():
        assert helpers.rel_err(p.grad, rg[k].grad) < 1e-5, k


@pytest.mark.skipif(not reference_loader.available(), reason='verbatim reference not on this host')
def test_oracle_jk_and_dense_tuple_match_reference():
    """jk=True (DenseJK) and the dense-tuple input format (network.py:254-256)."""
    graphs = helpers.make_graphs([30, 22], [5, 5], seed=9, messy=False)
    batch = graph.Batch.from_data_list(graphs)
    adj, counts = restatement.densify_edges(batch.edge_index, batch.batch)
    x = restatement.pad_nodes(batch.x, batch.batch, counts)
    args = helpers.encoder_args(120)
    kw = dict(concat=True, gcn_name='SAGE', load_data_sparse=False, jk=True)
    torch.manual_seed(2)
    ref = reference_loader.build_reference_encoder(*args, **kw)
    orc = restatement.OracleEncoder(*args, **kw)
    orc.load_state_dict(ref.state_dict())
    ref.eval(), orc.eval()
    with torch.no_grad():
        a = reference_loader.run_reference(ref, (x, adj.clone(), counts))
        b = orc((x, adj.clone(), counts))
    assert helpers.rel_err(b, a) < 1e-6


def test_shim_matches_restated_conv():
    torch.manual_seed(0)
    a = pyg_shim.DenseSAGEConv(16, 20)
    b = restatement.OracleSAGEConv(16, 20)
    b.load_state_dict(a.state_dict())
    x = torch.randn(2, 9, 16)
    adj = (torch.rand(2, 9, 9) > 0.6).float()
    mask = torch.ones(2, 9, 1)
    mask[1, 6:] = 0
    assert torch.equal(a(x, adj, mask, False), b(x, adj, mask, False))


def test_jk_reference_formulas_match_autograd():
    """The per-operator DenseJK formulas (forward, hand-written backward, weight-gradient contractions) of
    oracle/op_reference.py against torch.nn.LSTM + autograd in float64 (DenseJK.forward, network.py:36-52)."""
    from oracle import op_reference
    ref = op_reference.TorchBackend()
    torch.manual_seed(0)
    c, layers, m = 20, 3, 37
    hid = c * layers // 2
    lstm = torch.nn.LSTM(c, hid, bidirectional=True, batch_first=True).double()
    att = torch.nn.Linear(2 * hid, 1).double()
    feats = torch.randn(m, layers * c, dtype=torch.double, requires_grad=True)
    seq = feats.view(m, layers, c)
    score, _ = lstm(seq)
    alpha = torch.softmax(att(score).squeeze(-1), -1)
    y = (seq * alpha.unsqueeze(-1)).sum(1)
    g = torch.randn(m, c, dtype=torch.double)
    y.backward(g)
    names = ['weight_ih_l0', 'weight_hh_l0', 'bias_ih_l0', 'bias_hh_l0']
    params = [getattr(lstm, n).detach() for n in names] + [getattr(lstm, n + '_reverse').detach() for n in names]
    out = torch.zeros(m, c, dtype=torch.double)
    tape = torch.zeros(ref.jk_tape_floats(m, c, layers), dtype=torch.double)
    ref.jk_lstm_fwd(feats.detach(), c, layers, params, att.weight.detach(), att.bias.detach(), out, tape)
    assert float((out - y.detach()).abs().max()) < 1e-12
    gf = torch.zeros(m, layers * c, dtype=torch.double)
    dg = torch.zeros(2, layers, 4 * hid, m, dtype=torch.double)
    dsc = torch.zeros(layers, m, dtype=torch.double)
    ref.jk_lstm_bwd(g, feats.detach(), c, layers, params, att.weight.detach(), tape, gf, dg, dsc)
    assert float((gf - feats.grad).abs().max()) < 1e-12
    _, hs, _ = ref._jk_tape_views(tape, m, c, layers)
    xs = [feats.detach()[:, t * c:(t + 1) * c] for t in range(layers)]
    for d, suf in enumerate(['', '_reverse']):
        order = list(range(layers)) if d == 0 else list(range(layers - 1, -1, -1))
        d_wih = sum(dg[d, t] @ xs[t] for t in range(layers))
        d_whh = sum(dg[d, order[p]] @ hs[d, order[p - 1]].t() for p in range(1, layers))
        d_b = sum(dg[d, t].sum(1) for t in range(layers))
        assert float((d_wih - getattr(lstm, 'weight_ih_l0' + suf).grad).abs().max()) < 1e-12
        assert float((d_whh - getattr(lstm, 'weight_hh_l0' + suf).grad).abs().max()) < 1e-12
        assert float((d_b - getattr(lstm, 'bias_hh_l0' + suf).grad).abs().max()) < 1e-12
    d_aw = torch.cat([sum(hs[d, t] @ dsc[t] for t in range(layers)) for d in range(2)])
    assert float((d_aw - att.weight.grad.view(-1)).abs().max()) < 1e-12
    assert float((dsc.sum() - att.bias.grad).abs().max()) < 1e-12


def test_ell16_reference_table_invariants():
    """Property test (hypothesis) of the ELL16 restatement the GPU builder is compared with bit-exactly: for random
    CSRs the table is a re-ordering of every row's entries (weights stay attached to their columns), pad slots
    follow the valid ones, overflow rows are flagged, and the position map permutes each block of 8 rows."""
    from hypothesis import given, settings, strategies as st
    from oracle import op_reference
    ref = op_reference.TorchBackend()

    @settings(max_examples=30, deadline=None)
    @given(st.integers(1, 3), st.integers(1, 41), st.sampled_from([8, 16]), st.integers(0, 20), st.integers(0, 1),
           st.integers(0, 2 ** 31 - 1))
    def check(b, nmax, width, maxdeg, pair, seed):
        g = torch.Generator().manual_seed(seed)
        rows = b * nmax
        deg = torch.randint(0, min(maxdeg, nmax) + 1, (rows,), generator=g)
        rowptr = torch.zeros(rows + 1, dtype=torch.int32)
        rowptr[1:] = deg.cumsum(0).to(torch.int32)
        cols = [torch.randperm(nmax, generator=g)[:int(deg[r])].sort().values + (r // nmax) * nmax for r in range(rows)]
        col = torch.cat(cols + [torch.zeros(1, dtype=torch.long)]).to(torch.int32)
        val = torch.rand(int(rowptr[-1]) + 1, generator=g)
        trows = ref.ell16_rows(b, nmax)
        ell = torch.zeros(trows * width, dtype=torch.int16)
        ellw, ells = torch.zeros(trows * width), torch.zeros(trows)
        perm = torch.zeros(trows // 8, dtype=torch.int32)
        ref.csr_to_ell16(rowptr, col, val, None, b, nmax, width, pair, ell, ellw, ells, perm)
        table = ell.view(trows, width).long() & 0xFFFF
        tags = (table[:, 0] >> 12) & 7                    # slot 0 carries the served local row in bits 12-14
        table[:, 0] &= 0xFFF
        wt = ellw.view(trows, width)
        nrow8 = trows // b
        seen = set()
        for t in range(trows):
            gidx, p = divmod(t, nrow8)
            pw = int(perm[t // 8]) & 0xFFFFFF
            lrow = p // 8 * 8 + ((pw >> (3 * (p % 8))) & 7)
            assert int(tags[t]) == (pw >> (3 * (p % 8))) & 7
            seen.add((gidx, lrow))
            if lrow >= nmax:
                assert bool((table[t] == nmax).all())
                continue
            r = gidx * nmax + lrow
            lo, hi = int(rowptr[r]), int(rowptr[r + 1])
            if hi - lo > width:
                assert int(table[t, 0]) == nmax + 1
                continue
            n = hi - lo
            assert bool((table[t, n:] == nmax).all()) and bool((wt[t, n:] == 0).all())       # pads last, weight 0
            got = sorted(zip(table[t, :n].tolist(), wt[t, :n].tolist()))
            want = sorted(zip((col[lo:hi].long() - gidx * nmax).tolist(), val[lo:hi].tolist()))
            assert got == want
            if not pair:
                assert table[t, :n].tolist() == (col[lo:hi].long() - gidx * nmax).tolist() and lrow == p
        assert len(seen) == trows                                                            # a permutation per block

    check()


@pytest.mark.skipif(not os.path.isfile('/root/reference/common/utils.py'), reason='reference checkout not mounted')
def test_farthest_sampler_restatement_matches_reference_source():
    """The oracle's farthest-point sampler (op_reference.farthest_point_sample) against FarthestSampler of the
    reference itself (common/utils.py:187-197), EXECUTED from the reference source (imported behind the
    torch_geometric stand-ins) on the int16 distance matrix of construct_feature_graph.py:17-24."""
    import sys
    import numpy as np
    from cgc_net_b200 import compat
    from oracle import op_reference
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k.split('.')[0] in ('torch_geometric', 'common')}
    for k in saved:
        sys.modules.pop(k, None)
    compat.install(force=True)
    sys.path.insert(0, '/root/reference')
    try:
        from common.utils import FarthestSampler            # the reference class, unmodified
        ref = op_reference.TorchBackend()
        g = torch.Generator().manual_seed(3)
        for n, k in ((57, 20), (400, 280), (1, 1), (33, 33)):
            pos = torch.rand(n, 2, generator=g) * 900.0
            p = pos.double().numpy()
            arr = np.sqrt((p[:, 0, np.newaxis].T - p[:, 0, np.newaxis]) ** 2 +
                          (p[:, 1, np.newaxis].T - p[:, 1, np.newaxis]) ** 2).astype(np.int16)
            np.random.seed(n)
            want = FarthestSampler()(arr, k)                 # draws its first index from numpy's global stream
            np.random.seed(n)
            first = np.random.randint(arr.shape[0])
            assert int(want[0]) == first
            out = torch.zeros(k, dtype=torch.int32)
            ref.farthest_point_sample(pos, torch.tensor([0, n], dtype=torch.int32), torch.tensor([first], dtype=torch.int32),
                                      torch.tensor([0, k], dtype=torch.int32), n, out)
            assert out.tolist() == want.tolist()
    finally:
        sys.path.remove('/root/reference')
        for k in list(sys.modules):
            if k.split('.')[0] in ('torch_geometric', 'common'):
                sys.modules.pop(k, None)
        sys.modules.update({k: v for k, v in saved.items() if v is not None})
