"""The module surface and the hand-written backward of cgc_net_b200/ops.py, checked on CPU by
substituting the per-operator formulas of oracle/op_reference.py for the CUDA library."""
import os

import numpy as np
import pytest
import torch

import helpers
from cgc_net_b200 import _backend, graph, network, ops
from oracle import restatement
from test_oracle import GOLDEN_ENCODERS, _batch_from, _load, _variant_kwargs


def _grad_check(ours, orc, batch):
    """Gradients against the float64 oracle: 1e-4 relative, or -- for the cancellation-residue tensors whose absolute
    error is pure fp32 summation-order noise -- within 10x the fp32 oracle's own error (same bound as the GPU test's
    exact-fp32 mode)."""
    _, _, truth = helpers.oracle_fp64(orc, batch)
    helpers.assert_grads_close(ours, truth, {k: p.grad for k, p in orc.named_parameters()}, tol=1e-4, slack=10.0,
                               floor=2e-3)


@pytest.mark.parametrize('name', GOLDEN_ENCODERS)
def test_module_matches_golden(torch_backend, name):
    g = _load(name)
    args = (int(g['max_nodes']), 16, 20, 20, True, True, 20, 3, float(g['ratio']), [50])
    m = network.SoftPoolingGcnEncoder(*args, concat=True, collect_assign=True, load_data_sparse=True,
                                      **_variant_kwargs(g))
    m.load_state_dict({k[len('param/'):]: torch.from_numpy(g[k]) for k in g.files if k.startswith('param/')})
    m.train()
    logits, loss = m(_batch_from(g))
    assert helpers.rel_err(logits, torch.from_numpy(g['logits'])) < 1e-4
    assert abs(float(loss.detach()) - float(g['loss'])) < 1e-5
    loss.backward()
    for k, p in m.named_parameters():
        ref = g['grad/' + k]
        assert np.abs(p.grad.numpy() - ref).max() <= 3e-4 * max(np.abs(ref).max(), 1e-3), k
    for k, v in m.state_dict().items():
        if 'running_' in k:
            np.testing.assert_allclose(v.numpy(), g['after/' + k], rtol=1e-4, atol=1e-6)
    for i, idx in enumerate(m.assign_index):
        safe = g['assign_top2_margin/%d' % i] > 1e-6
        assert np.array_equal(idx.numpy()[safe], g['assign_argmax/%d' % i][safe])     # bit-exact indices
    m.eval()
    with torch.no_grad():
        assert helpers.rel_err(m(_batch_from(g)), torch.from_numpy(g['eval_logits'])) < 1e-4


@pytest.mark.parametrize('norm_adj', [False, True])
def test_module_matches_oracle_ragged(torch_backend, norm_adj):
    graphs = helpers.make_graphs([33, 80, 57, 12], [5, 8, 12, 3], seed=4)
    batch = graph.Batch.from_data_list(graphs)
    ours, orc = helpers.build_pair(helpers.encoder_args(300), load_data_sparse=True, norm_adj=norm_adj)
    ours.train(), orc.train()
    lo, losso = ours(batch)
    lr, lossr = orc(batch)
    assert helpers.rel_err(lo, lr) < 1e-4
    losso.backward(), lossr.backward()
    _grad_check(ours, orc, batch)
    for (k, a), (_, b) in zip(sorted(ours.named_buffers()), sorted(orc.named_buffers())):
        assert helpers.rel_err(a.float(), b.float()) < 1e-4, k


def test_batch_without_host_counts(torch_backend):
    """A PyG-style batch carrying only x / edge_index / batch / y (no host metadata)."""
    graphs = helpers.make_graphs([20, 31], [4, 6], seed=8)
    b = graph.Batch.from_data_list(graphs)
    plain = graph.Batch(b.x, b.edge_index, b.batch, b.y, None)
    ours, orc = helpers.build_pair(helpers.encoder_args(200), load_data_sparse=True)
    ours.eval(), orc.eval()
    with torch.no_grad():
        assert helpers.rel_err(ours(plain), orc(plain)) < 1e-4


@pytest.mark.parametrize('norm_adj', [False, True])
def test_dense_tuple_input(torch_backend, norm_adj):
    """(x, adj, num_nodes[, label]) input of network.py:254-256."""
    graphs = helpers.make_graphs([25, 40], [5, 6], seed=6)
    b = graph.Batch.from_data_list(graphs)
    adj, counts = restatement.densify_edges(b.edge_index, b.batch)
    x = restatement.pad_nodes(b.x, b.batch, counts)
    ours, orc = helpers.build_pair(helpers.encoder_args(200), load_data_sparse=False, norm_adj=norm_adj)
    ours.train(), orc.train()
    lo, losso = ours((x, adj.clone(), counts, b.y))
    lr, lossr = orc((x, adj.clone(), counts, b.y))
    assert helpers.rel_err(lo, lr) < 1e-4
    losso.backward(), lossr.backward()
    _grad_check(ours, orc, (x, adj.clone(), counts, b.y))


def test_state_dict_names_and_size():
    """Checkpoint compatibility: BASELINE model has 1 410 275 parameters."""
    m = network.SoftPoolingGcnEncoder(11404, 16, 20, 20, True, True, 20, 3, 0.1, [50], concat=True)
    o = restatement.OracleEncoder(11404, 16, 20, 20, True, True, 20, 3, 0.1, [50], concat=True)
    assert sum(p.numel() for p in m.parameters()) == 1410275
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == \
           {k: tuple(v.shape) for k, v in o.state_dict().items()}
    assert tuple(m.GCN_pool_1.gcn3.weight.shape) == (20, 1140)
    assert tuple(m.GCN_pool_1.lin.weight.shape) == (1140, 1180)
    o.load_state_dict(m.state_dict())


@pytest.mark.parametrize('norm_adj', [False, True])
def test_jumping_knowledge_matches_oracle(torch_backend, norm_adj):
    """jk=True (DenseJK, the published configuration): changes the stage widths to 20 / 60 (network.py:150-167)."""
    graphs = helpers.make_graphs([41, 67, 30], [5, 8, 6], seed=13)
    batch = graph.Batch.from_data_list(graphs)
    ours, orc = helpers.build_pair(helpers.encoder_args(300), load_data_sparse=True, norm_adj=norm_adj, jk=True)
    assert {k: tuple(v.shape) for k, v in ours.state_dict().items()} == \
           {k: tuple(v.shape) for k, v in orc.state_dict().items()}
    ours.train(), orc.train()
    lo, losso = ours(batch)
    lr, lossr = orc(batch)
    assert helpers.rel_err(lo, lr) < 1e-4
    losso.backward(), lossr.backward()
    _grad_check(ours, orc, batch)


@pytest.mark.parametrize('gcn_name,activation,norm_adj', [('GIN', 'relu', False), ('GIN', 'elu', True),
                                                           ('SAGE', 'elu', False), ('SAGE', 'leakyrelu', True),
                                                           ('GIN', 'leakyrelu', False)])
def test_gin_and_activations_match_oracle(torch_backend, gcn_name, activation, norm_adj):
    """DenseGINConv variant (network.py:96-99) and the elu / leakyrelu activations (network.py:82-90)."""
    graphs = helpers.make_graphs([41, 67, 30], [5, 8, 6], seed=21)
    batch = graph.Batch.from_data_list(graphs)
    ours, orc = helpers.build_pair(helpers.encoder_args(300), load_data_sparse=True, norm_adj=norm_adj,
                                   gcn_name=gcn_name, activation=activation)
    assert {k: tuple(v.shape) for k, v in ours.state_dict().items()} == \
           {k: tuple(v.shape) for k, v in orc.state_dict().items()}
    ours.train(), orc.train()
    lo, losso = ours(batch)
    lr, lossr = orc(batch)
    assert helpers.rel_err(lo, lr) < 1e-4
    losso.backward(), lossr.backward()
    _grad_check(ours, orc, batch)
    for (k, a), (_, b) in zip(sorted(ours.named_buffers()), sorted(orc.named_buffers())):
        assert helpers.rel_err(a.float(), b.float()) < 1e-4, k
    ours.eval(), orc.eval()
    with torch.no_grad():
        assert helpers.rel_err(ours(batch), orc(batch)) < 1e-4


def test_gin_dense_tuple_input(torch_backend):
    graphs = helpers.make_graphs([25, 40], [5, 6], seed=6)
    b = graph.Batch.from_data_list(graphs)
    adj, counts = restatement.densify_edges(b.edge_index, b.batch)
    x = restatement.pad_nodes(b.x, b.batch, counts)
    ours, orc = helpers.build_pair(helpers.encoder_args(200), load_data_sparse=False, gcn_name='GIN')
    ours.train(), orc.train()
    lo, losso = ours((x, adj.clone(), counts, b.y))
    lr, lossr = orc((x, adj.clone(), counts, b.y))
    assert helpers.rel_err(lo, lr) < 1e-4
    losso.backward(), lossr.backward()
    _grad_check(ours, orc, (x, adj.clone(), counts, b.y))


def test_unknown_activation_raises():
    with pytest.raises(ValueError):
        network.SoftPoolingGcnEncoder(100, 16, 20, 20, True, True, 20, 3, activation='gelu')


def test_graph_build_matches_to_dense_adj(torch_backend):
    """CSR/CSC against the dense adjacency of model/utils.py:3-37 (duplicates, isolated rows)."""
    graphs = helpers.make_graphs([17, 9, 30], [4, 3, 7], seed=2)
    b = graph.Batch.from_data_list(graphs)
    adj, counts = restatement.densify_edges(b.edge_index, b.batch)
    xp, g = ops.build_sparse_graph(b.x, b.edge_index, b.batch, b.num_nodes_list, False)
    rows = g.rows
    dense = torch.zeros(rows, rows)
    nnz = int(g.rowptr[-1])
    deg = (g.rowptr[1:] - g.rowptr[:-1]).long()
    r = torch.repeat_interleave(torch.arange(rows), deg)
    dense[r, g.col[:nnz].long()] = 1
    blocks = torch.block_diag(*[adj[i] for i in range(adj.shape[0])])
    assert torch.equal(dense, blocks)
    assert int(adj.sum()) == nnz
    assert torch.equal(xp.view(adj.shape[0], g.nmax, -1), restatement.pad_nodes(b.x, b.batch, counts))
    # CSC is the exact transpose
    dt = torch.zeros(rows, rows)
    cdeg = (g.colptr[1:] - g.colptr[:-1]).long()
    c = torch.repeat_interleave(torch.arange(rows), cdeg)
    dt[g.row[:nnz].long(), c] = 1
    assert torch.equal(dt, dense)


def test_node_balanced_split():
    graphs = [graph.Data(x=torch.zeros(n, 1)) for n in (10, 10, 10, 10, 40, 5, 5, 30)]
    parts = graph.split_by_node_count(graphs, 2)
    assert sum(len(p) for p in parts) == len(graphs)
    assert [d.num_nodes for d in parts[0]] == [10, 10, 10, 10]


def test_product_path_has_no_cpu_fallback():
    """CPU tensors must be rejected by the CUDA backend wrapper (and the library must exist)."""
    _backend.use_backend_for_testing(None)
    try:
        be = _backend.get()
    except _backend.CgcError:
        pytest.skip('library not built')
    with pytest.raises(_backend.CgcError):
        be.row_inv_degree(torch.zeros(3, dtype=torch.int32), torch.zeros(2))


def test_property_graph_build_random(torch_backend):
    """Property test (hypothesis): random sizes / degrees / duplicate and unsorted edges -- the CSR of the
    padded rows always equals the reference's dense adjacency and the CSC is its exact transpose."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=25, deadline=None)
    @given(st.lists(st.integers(1, 40), min_size=1, max_size=4), st.integers(0, 12), st.integers(0, 2 ** 31 - 1))
    def check(sizes, deg, seed):
        g = torch.Generator().manual_seed(seed)
        datas = []
        for n in sizes:
            e = n * deg
            ei = torch.randint(0, n, (2, e), generator=g) if e else torch.zeros(2, 0, dtype=torch.long)
            datas.append(graph.Data(x=torch.randn(n, 4, generator=g), edge_index=ei, y=torch.zeros(1, dtype=torch.long)))
        b = graph.Batch.from_data_list(datas)
        adj, counts = restatement.densify_edges(b.edge_index, b.batch) if b.edge_index.shape[1] else \
            (torch.zeros(len(sizes), max(sizes), max(sizes)), torch.tensor(sizes))
        xp, gr = ops.build_sparse_graph(b.x, b.edge_index, b.batch, b.num_nodes_list, False)
        rows = gr.rows
        nnz = int(gr.rowptr[-1])
        dense = torch.zeros(rows, rows)
        deg_r = (gr.rowptr[1:] - gr.rowptr[:-1]).long()
        dense[torch.repeat_interleave(torch.arange(rows), deg_r), gr.col[:nnz].long()] = 1
        assert torch.equal(dense, torch.block_diag(*[adj[i] for i in range(adj.shape[0])]))
        dt = torch.zeros(rows, rows)
        deg_c = (gr.colptr[1:] - gr.colptr[:-1]).long()
        dt[gr.row[:nnz].long(), torch.repeat_interleave(torch.arange(rows), deg_c)] = 1
        assert torch.equal(dt, dense)
        assert torch.equal(gr.rs, 1.0 / deg_r.float().clamp(min=1))

    check()


def test_radius_graph_wrapper_feeds_the_encoder(torch_backend):
    """graph.radius_graph (reference call: radius_graph(pos, 100, None, True, 8), dataflow/data.py:348) produces
    an edge_index the encoder consumes: neighbour -> centre pairs, self loop first, <= k per centre."""
    g = torch.Generator().manual_seed(5)
    sizes = [60, 35]
    pos = torch.rand(sum(sizes), 2, generator=g) * 300.0
    ei = graph.radius_graph(pos, 100.0, None, True, 8, num_nodes_list=sizes)
    assert ei.dtype == torch.long and ei.shape[0] == 2
    indeg = torch.bincount(ei[1], minlength=sum(sizes))
    assert int(indeg.max()) <= 8 and int(indeg.min()) >= 1
    d = (pos[ei[0]] - pos[ei[1]]).norm(dim=1)
    assert float(d.max()) < 100.0
    same_graph = (ei[0] < 60) == (ei[1] < 60)
    assert bool(same_graph.all())
    # brute force: every centre's list equals its k nearest within r
    dm = torch.cdist(pos.double(), pos.double())
    dm[:60, 60:] = float('inf')
    dm[60:, :60] = float('inf')
    for c in (0, 17, 59, 60, 94):
        want = [int(j) for j in dm[c].argsort()[:8] if dm[c, j] < 100.0]
        assert ei[0][ei[1] == c].tolist() == want
    batch = graph.Batch(torch.randn(sum(sizes), 16, generator=g), ei,
                        torch.cat([torch.full((n,), i) for i, n in enumerate(sizes)]), torch.tensor([0, 2]), sizes)
    ours, orc = helpers.build_pair(helpers.encoder_args(200), load_data_sparse=True)
    ours.eval(), orc.eval()
    with torch.no_grad():
        assert helpers.rel_err(ours(batch), orc(batch)) < 1e-4


def test_compat_data_containers_and_loaders():
    """cgc_net_b200.compat: the torch_geometric 1.2.1 surface the reference's loaders use (dataflow/data.py:4-11)."""
    from cgc_net_b200 import compat
    d = compat.Data(x=torch.randn(5, 3), edge_index=torch.tensor([[0, 1], [1, 2]]), y=torch.tensor([1]),
                    pos=torch.rand(5, 2), patch_idx=torch.tensor([7]))
    assert d.num_nodes == 5 and set(k for k, _ in d) >= {'x', 'edge_index', 'y', 'pos', 'patch_idx'}
    choice = torch.tensor([4, 0, 2])
    for key, item in d:                                   # the reference's node subsampling idiom (data.py:241-243)
        if torch.is_tensor(item) and item.size(0) == 5:
            d[key] = item[choice]
    assert d.x.shape == (3, 3) and d.pos.shape == (3, 2)

    class DS(compat.Dataset):
        def __init__(self):
            super(DS, self).__init__(root=None)

        def __len__(self):
            return 4

        def get(self, idx):
            g = graph.synthetic_cell_graph(6 + idx, 3, 4, seed=idx)
            return compat.Data(x=g.x, edge_index=g.edge_index, y=g.y, pos=g.pos, patch_idx=torch.tensor([idx]))

    lists = list(compat.DataListLoader(DS(), batch_size=2))
    assert len(lists) == 2 and isinstance(lists[0], list) and lists[0][1].num_nodes == 7
    batches = list(compat.DataLoader(DS(), batch_size=4))
    b = batches[0]
    assert b.num_graphs == 4 and b.x.shape[0] == 6 + 7 + 8 + 9 and b.patch_idx.tolist() == [0, 1, 2, 3]
    assert int(b.edge_index.max()) < b.x.shape[0] and b.batch.tolist() == sorted(b.batch.tolist())
    # utils
    ei = torch.tensor([[0, 2, 2], [1, 0, 2]])
    adj = compat.sparse_to_dense(ei)
    assert adj.tolist() == [[0, 1, 0], [0, 0, 0], [1, 0, 1]]
    back, _ = compat.dense_to_sparse(adj)
    assert sorted(map(tuple, back.t().tolist())) == sorted(map(tuple, ei.t().tolist()))
    dense, counts = compat.to_dense_batch(torch.arange(5.0).view(5, 1), torch.tensor([0, 0, 1, 1, 1]))
    assert counts.tolist() == [2, 3] and dense.shape == (2, 3, 1) and float(dense[0, 2, 0]) == 0.0


def test_compat_radius_graph_cpu_matches_kernel_semantics(torch_backend):
    """CPU path of compat.radius_graph (cKDTree, as torch_cluster does) == the per-operator formula of the GPU
    kernel; flow argument; loop=False drops the centre."""
    from cgc_net_b200 import compat
    g = torch.Generator().manual_seed(3)
    pos = torch.rand(50, 2, generator=g) * 200.0
    batch = torch.cat([torch.zeros(30, dtype=torch.long), torch.ones(20, dtype=torch.long)])
    a = compat.radius_graph(pos, 60.0, batch, True, 5)
    b = graph.radius_graph(pos, 60.0, batch, True, 5)                 # library path (CPU formulas injected)
    assert torch.equal(a, b)
    t2s = compat.radius_graph(pos, 60.0, batch, True, 5, flow='target_to_source')
    assert torch.equal(t2s, a.flip(0))
    noloop = compat.radius_graph(pos, 60.0, batch, False, 5)
    assert bool((noloop[0] != noloop[1]).all())


@pytest.mark.parametrize('method', ['random', 'farthest', 'fuse'])
def test_sample_nodes_counts_and_uniqueness(torch_backend, method):
    """graph.sample_nodes (dataflow/data.py:195-225): int(n * ratio) distinct local ids per graph; the farthest-point
    prefix is reproducible from the generator seed."""
    g = torch.Generator().manual_seed(11)
    sizes = [40, 25, 3]
    pos = torch.rand(sum(sizes), 2, generator=g) * 500.0
    picks = graph.sample_nodes(pos, sizes, 0.5, method, generator=torch.Generator().manual_seed(5))
    again = graph.sample_nodes(pos, sizes, 0.5, method, generator=torch.Generator().manual_seed(5))
    for n, ids, ids2 in zip(sizes, picks, again):
        assert ids.numel() == int(n * 0.5) and ids.unique().numel() == ids.numel()
        assert ids.numel() == 0 or (int(ids.min()) >= 0 and int(ids.max()) < n)
        assert torch.equal(ids, ids2)


def test_eighteen_feature_input(torch_backend):
    """The reference's own feature width (16 appearance + 2 coordinate features, dataflow/data.py:22-53): not a
    multiple of 4, so layer 1 takes the generic gather while the ELL16 table still serves the wide A*S."""
    graphs = helpers.make_graphs([41, 67, 30], [5, 8, 6], feat=18, seed=31)
    batch = graph.Batch.from_data_list(graphs)
    args = (300, 18, 20, 20, True, True, 20, 3, 0.1, [50])
    ours, orc = helpers.build_pair(args, load_data_sparse=True)
    ours.train(), orc.train()
    lo, losso = ours(batch)
    lr, lossr = orc(batch)
    assert helpers.rel_err(lo, lr) < 1e-4
    losso.backward(), lossr.backward()
    _grad_check(ours, orc, batch)


def test_folded_batchnorm_backward_is_a_pure_reformulation(torch_backend, monkeypatch):
    """ops.FOLD_BN_BACKWARD routes the wide layers' backward through sage_dh_bn (BatchNorm backward formed inside the
    epilogue backward) instead of bn_bwd(apply) + sage_dh: both wirings must give the same gradients."""
    graphs = [graph.synthetic_cell_graph(n, 6, 16, seed=i) for i, n in enumerate([300, 400, 250])]
    batch = graph.Batch.from_data_list(graphs)
    grads = {}
    for fold in (True, False):
        monkeypatch.setattr(ops, 'FOLD_BN_BACKWARD', fold)
        torch.manual_seed(0)
        m = network.SoftPoolingGcnEncoder(400, 16, 20, 20, True, True, 20, 3, 0.1, [50], concat=True,
                                          load_data_sparse=True).train()
        _, loss = m(batch)
        loss.backward()
        grads[fold] = {k: p.grad.clone() for k, p in m.named_parameters()}
    for k, g in grads[True].items():
        scale = max(float(grads[False][k].abs().max()), 1e-6)
        assert float((g - grads[False][k]).abs().max()) <= 1e-5 * scale + 1e-8, k
