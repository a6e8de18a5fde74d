"""Every C-ABI entry point on the GPU against the per-operator CPU formulas (oracle/op_reference.py).

Bit-exact for index work (CSR/CSC construction, arg-max, arg-indices); fp32 tolerance for
floating point (stated per test).  All calls go through the ctypes binding of libcgc_sm100.so."""
import numpy as np
import pytest
import torch

import helpers
from cgc_net_b200 import graph, ops
from cgc_net_b200.ops import RowMask
from oracle import op_reference

pytestmark = pytest.mark.gpu
REF = op_reference.TorchBackend()


def _cu(a):
    if torch.is_tensor(a):
        return a.cuda()
    if isinstance(a, RowMask):
        return RowMask(a.num_nodes.cuda(), a.nmax)
    return a


def both(be, name, args, outs, rtol=1e-5, atol=1e-6, exact=()):
    """Run ``name`` on CPU formulas and on the CUDA library; compare the tensors at indices ``outs``.
    Strided views are reproduced on the GPU by cloning the base storage."""
    gpu_args = []
    for a in args:
        if torch.is_tensor(a) and a._base is not None:
            base = a._base.cuda()
            gpu_args.append(base.as_strided(a.shape, a.stride(), a.storage_offset()))
        else:
            gpu_args.append(_cu(a))
    getattr(REF, name)(*args)
    getattr(be, name)(*gpu_args)
    torch.cuda.synchronize()
    for i in outs:
        got, want = gpu_args[i].cpu(), args[i]
        if i in exact or not want.is_floating_point():
            assert torch.equal(got, want), '%s: output %d differs' % (name, i)
        else:
            torch.testing.assert_close(got, want, rtol=rtol, atol=atol, msg=lambda m: '%s out %d: %s' % (name, i, m))
    return gpu_args


def _mask(b, nmax, seed=0):
    g = torch.Generator().manual_seed(seed)
    nn = torch.randint(max(1, nmax // 2), nmax + 1, (b,), generator=g, dtype=torch.int32)
    nn[0] = nmax
    return RowMask(nn, nmax)


def _random_csr(rows, maxdeg, seed=0, block=None):
    """CSR with sorted unique columns; ``block`` keeps columns inside the row's graph block."""
    g = torch.Generator().manual_seed(seed)
    deg = torch.randint(0, maxdeg + 1, (rows,), generator=g)
    rowptr = torch.zeros(rows + 1, dtype=torch.int32)
    rowptr[1:] = deg.cumsum(0).to(torch.int32)
    cols = []
    for r in range(rows):
        lo, hi = (0, rows) if block is None else ((r // block) * block, (r // block + 1) * block)
        c = torch.randperm(hi - lo, generator=g)[:int(deg[r])] + lo
        cols.append(c.sort().values)
    col = torch.cat(cols).to(torch.int32) if cols else torch.zeros(0, dtype=torch.int32)
    return rowptr, torch.cat([col, torch.zeros(3, dtype=torch.int32)])


# ------------------------------------------------------------------------------- graph build
@pytest.mark.parametrize('norm_adj', [False, True])
@pytest.mark.parametrize('with_counts', [True, False])
def test_graph_build_bit_exact(cuda_backend, norm_adj, with_counts):
    graphs = helpers.make_graphs([37, 120, 64, 5, 201], [5, 8, 12, 3, 8], seed=1)
    b = graph.Batch.from_data_list(graphs)
    counts = b.num_nodes_list if with_counts else None
    from cgc_net_b200 import _backend
    _backend.use_backend_for_testing(REF)
    xr, gr = ops.build_sparse_graph(b.x, b.edge_index, b.batch, counts, norm_adj)
    _backend.use_backend_for_testing(None)
    bc = b.to('cuda')
    xg, gg = ops.build_sparse_graph(bc.x, bc.edge_index, bc.batch, counts, norm_adj)
    torch.cuda.synchronize()
    assert (gg.num_graphs, gg.nmax, gg.rows) == (gr.num_graphs, gr.nmax, gr.rows)
    assert torch.equal(xg.cpu(), xr)
    assert torch.equal(gg.num_nodes.cpu(), gr.num_nodes)
    nnz = int(gr.rowptr[-1])
    assert torch.equal(gg.rowptr.cpu(), gr.rowptr) and torch.equal(gg.col[:nnz].cpu(), gr.col[:nnz])
    assert torch.equal(gg.colptr.cpu(), gr.colptr) and torch.equal(gg.row[:nnz].cpu(), gr.row[:nnz])
    torch.testing.assert_close(gg.rs.cpu(), gr.rs, rtol=1e-6, atol=0)
    if norm_adj:
        torch.testing.assert_close(gg.val[:nnz].cpu(), gr.val[:nnz], rtol=1e-6, atol=0)
        torch.testing.assert_close(gg.val_t[:nnz].cpu(), gr.val_t[:nnz], rtol=1e-6, atol=0)


def test_graph_build_large_scan(cuda_backend):
    """> 1024 rows per scan tile and > 256 tiles: exercises the three-kernel scan."""
    graphs = graph.synthetic_batch(160, 1800, 8, 4, seed=3)
    b = graph.Batch.from_data_list(graphs).to('cuda')
    _, g = ops.build_sparse_graph(b.x, b.edge_index, b.batch, b.num_nodes_list, False)
    torch.cuda.synchronize()
    deg = (g.rowptr[1:] - g.rowptr[:-1]).cpu()
    assert int(g.rowptr[-1]) == 160 * 1800 * 8 and bool((deg == 8).all())
    cdeg = (g.colptr[1:] - g.colptr[:-1]).cpu()
    assert int(cdeg.sum()) == 160 * 1800 * 8
    col = g.col[:160 * 1800 * 8].view(-1, 8).cpu()
    assert bool((col[:, 1:] > col[:, :-1]).all())                 # sorted, unique
    assert bool(((col // 1800) == torch.arange(160 * 1800).view(-1, 1) // 1800).all())   # block diagonal


# ------------------------------------------------------------------------------- sparse gather
@pytest.mark.parametrize('feat', [16, 20, 60, 7, 1140])
@pytest.mark.parametrize('weighted', [False, True])
def test_spmm(cuda_backend, feat, weighted):
    rows = 300
    rowptr, col = _random_csr(rows, 9, seed=feat)
    nnz = int(rowptr[-1])
    g = torch.Generator().manual_seed(5)
    x = torch.randn(rows, feat, generator=g)
    val = torch.rand(nnz + 3, generator=g) if weighted else None
    ss = torch.rand(rows, generator=g) if weighted else None
    rsc = torch.rand(rows, generator=g)
    out = torch.zeros(rows, feat)
    both(cuda_backend, 'spmm', [rowptr, col, val, x, ss, rsc, out, False], [6], rtol=1e-5, atol=1e-5)
    out2 = torch.randn(rows, feat, generator=g)
    both(cuda_backend, 'spmm', [rowptr, col, val, x, None, None, out2, True], [6], rtol=1e-5, atol=1e-5)


def test_spmm_strided_slices(cuda_backend):
    """x and out as column slices of wider matrices (the concat buffer of GNN_Module)."""
    rows = 257
    rowptr, col = _random_csr(rows, 6, seed=2)
    g = torch.Generator().manual_seed(6)
    xw = torch.randn(rows, 60, generator=g)
    ow = torch.randn(rows, 60, generator=g)
    x, out = xw[:, 20:40], ow[:, 0:20]
    args = both(cuda_backend, 'spmm', [rowptr, col, None, x, None, None, out, True], [6], atol=1e-5)
    assert torch.equal(args[6]._base[:, 20:].cpu(), ow[:, 20:])      # nothing outside the slice touched


@pytest.mark.parametrize('feat,b,nmax', [(16, 3, 70), (20, 5, 333), (16, 64, 1800), (8, 150, 40), (32, 2, 1500)])
@pytest.mark.parametrize('weighted', [False, True])
def test_sage_gather_tile(cuda_backend, feat, b, nmax, weighted):
    """TMA-staged shared-memory gather == generic gather (same sums, same order)."""
    rows = b * nmax
    rowptr, col = _random_csr(rows, 8, seed=nmax, block=nmax) if rows < 5000 else _knn_csr(b, nmax)
    nnz = int(rowptr[-1])
    g = torch.Generator().manual_seed(7)
    x = torch.randn(rows, feat, generator=g)
    val = torch.rand(nnz + 3, generator=g) if weighted else None
    rsc = torch.rand(rows, generator=g) if weighted else None
    out = torch.zeros(rows, feat)
    both(cuda_backend, 'sage_gather_mean_fwd', [rowptr, col, val, rsc, x, b, nmax, out], [7], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('width,maxdeg', [(8, 8), (8, 11), (16, 16), (16, 19)])
@pytest.mark.parametrize('pair', [0, 1])
@pytest.mark.parametrize('weighted', [False, True])
def test_csr_to_ell16_bit_exact(cuda_backend, width, maxdeg, pair, weighted):
    """Row-padded CSR table: indices, weights and the position map bit-exact (pad slots, overflow marks,
    slot-0 choice, parity slot order, in-block row pairing); nmax not a multiple of 8."""
    b, nmax = 3, 57
    rows = b * nmax
    rowptr, col = _random_csr(rows, maxdeg, seed=width + maxdeg, block=nmax)
    nnz = int(rowptr[-1])
    val = torch.rand(nnz + 3, generator=torch.Generator().manual_seed(1)) if weighted else None
    trows = REF.ell16_rows(b, nmax)
    assert trows == cuda_backend.ell16_rows(b, nmax) == 3 * 64
    ell = torch.zeros(trows * width, dtype=torch.int16)
    ellw = torch.zeros(trows * width) if weighted else None
    ells = torch.zeros(trows)
    perm = torch.zeros(trows // 8, dtype=torch.int32)
    rsc = torch.rand(rows, generator=torch.Generator().manual_seed(2)) if weighted else None
    both(cuda_backend, 'csr_to_ell16', [rowptr, col, val, rsc, b, nmax, width, pair, ell, ellw, ells, perm],
         [8, 9, 10, 11] if weighted else [8, 10, 11], exact=(8, 9, 10, 11))
    # every row's entry set survives the re-ordering
    table = (ell.view(trows, width).long() & 0xFFFF)
    assert torch.equal((table[:, 0] >> 12) & 7, (perm.long()[torch.arange(trows) // 8] >> (3 * (torch.arange(trows) % 8))) & 7)
    table[:, 0] &= 0xFFF                                  # slot 0 carries the served local row in bits 12-14
    pos = torch.arange(trows)
    lrow = (pos % 64) // 8 * 8 + ((perm.long()[pos // 8] >> (3 * (pos % 8))) & 7)
    deg = (rowptr[1:] - rowptr[:-1]).long()
    for t in range(trows):
        if int(lrow[t]) >= nmax:
            assert bool((table[t] == nmax).all())
            continue
        r = (t // 64) * nmax + int(lrow[t])
        if int(deg[r]) > width:
            assert int(table[t, 0]) == nmax + 1
        else:
            want = sorted((col[int(rowptr[r]):int(rowptr[r + 1])].long() - (t // 64) * nmax).tolist())
            assert sorted(v for v in table[t].tolist() if v != nmax) == want


@pytest.mark.parametrize('feat,b,nmax,width,maxdeg', [(16, 3, 70, 8, 8), (16, 5, 333, 8, 10), (16, 64, 1800, 8, 8),
                                                      (8, 150, 40, 16, 16), (32, 2, 1500, 16, 18),
                                                      (20, 4, 1100, 8, 8), (16, 2, 2999, 8, 8)])
@pytest.mark.parametrize('weighted', [False, True])
def test_sage_gather_ell(cuda_backend, feat, b, nmax, width, maxdeg, weighted):
    """Fixed-pitch TMA-staged gather == CPU formulas == generic CSR gather (fp32 tolerance: slot order differs)."""
    rows = b * nmax
    rowptr, col = _random_csr(rows, maxdeg, seed=nmax, block=nmax) if rows < 5000 else _knn_csr(b, nmax)
    nnz = int(rowptr[-1])
    g = torch.Generator().manual_seed(7)
    x = torch.randn(rows, feat, generator=g)
    val = torch.rand(nnz + 3, generator=g) if weighted else None
    rsc = torch.rand(rows, generator=g) if weighted else None
    trows = REF.ell16_rows(b, nmax)
    ell = torch.zeros(trows * width, dtype=torch.int16)
    ellw = torch.zeros(trows * width) if weighted else None
    ells = torch.zeros(trows)
    perm = torch.zeros(trows // 8, dtype=torch.int32)
    REF.csr_to_ell16(rowptr, col, val, rsc, b, nmax, width, int(feat == 16), ell, ellw, ells, perm)
    out = torch.zeros(rows, feat)
    both(cuda_backend, 'sage_gather_ell_fwd', [ell, ellw, ells, perm, width, rowptr, col, val, x, b, nmax, out], [11],
         rtol=1e-5, atol=1e-5)
    want = torch.zeros(rows, feat)
    REF.sage_gather_mean_fwd(rowptr, col, val, rsc, x, b, nmax, want)
    torch.testing.assert_close(out, want, rtol=1e-5, atol=1e-5)


def _knn_csr(b, nmax, k=8):
    graphs = graph.synthetic_batch(b, nmax, k, 1, seed=11)
    bt = graph.Batch.from_data_list(graphs)
    from cgc_net_b200 import _backend
    _backend.use_backend_for_testing(REF)
    _, g = ops.build_sparse_graph(bt.x, bt.edge_index, bt.batch, bt.num_nodes_list, False)
    _backend.use_backend_for_testing(None)
    return g.rowptr, torch.cat([g.col, torch.zeros(3, dtype=torch.int32)])


# ------------------------------------------------------------------------------- SAGE epilogue
@pytest.mark.parametrize('cin,cout', [(16, 20), (20, 20), (60, 20), (20, 114), (20, 1140), (60, 300)])
@pytest.mark.parametrize('masked', [False, True])
def test_sage_epilogue_and_dh(cuda_backend, cin, cout, masked):
    b, nmax = 3, 53
    rows = b * nmax
    g = torch.Generator().manual_seed(cin * 1000 + cout)
    sa = torch.randn(rows, cin, generator=g)
    sa[5] = 0                                            # a row whose pre-normalisation vector is just the bias
    w = torch.randn(cin, cout, generator=g) / cin ** 0.5
    bias = torch.randn(cout, generator=g) * 0.1
    mask = _mask(b, nmax) if masked else None
    y = torch.zeros(rows, cout)
    both(cuda_backend, 'sage_epilogue_fwd', [sa, w, bias, mask, y], [4], rtol=1e-5, atol=1e-6)
    gy = torch.randn(rows, cout, generator=g)
    dh = torch.zeros(rows, cout)
    both(cuda_backend, 'sage_dh', [gy, sa, w, bias, y, mask, dh], [6], rtol=1e-4, atol=1e-5)
    # in place (dh aliases gy), no bias, strided y
    yw = torch.zeros(rows, cout + 8)
    both(cuda_backend, 'sage_epilogue_fwd', [sa, w, None, mask, yw[:, 4:4 + cout]], [4], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('act', [1, 2])
@pytest.mark.parametrize('cin,cout', [(16, 20), (20, 300)])
def test_sage_epilogue_activations(cuda_backend, act, cin, cout):
    """elu / leakyrelu epilogues and their gradients (narrow, wide and fused-narrow kernels)."""
    b, nmax = 3, 53
    rows = b * nmax
    g = torch.Generator().manual_seed(act * 100 + cout)
    sa = torch.randn(rows, cin, generator=g)
    w = torch.randn(cin, cout, generator=g) / cin ** 0.5
    bias = torch.randn(cout, generator=g) * 0.1
    mask = _mask(b, nmax)
    y = torch.zeros(rows, cout)
    getattr(REF, 'sage_epilogue_fwd')(sa, w, bias, mask, y, act)
    yg = torch.zeros(rows, cout).cuda()
    cuda_backend.sage_epilogue_fwd(sa.cuda(), w.cuda(), bias.cuda(), _cu(mask), yg, act)
    torch.testing.assert_close(yg.cpu(), y, rtol=1e-5, atol=1e-6)
    assert bool((y < 0).any())                                   # the negative branch is exercised
    gy = torch.randn(rows, cout, generator=g)
    dh = torch.zeros(rows, cout)
    REF.sage_dh(gy, sa, w, bias, y, mask, dh, act)
    dhg = torch.zeros(rows, cout).cuda()
    cuda_backend.sage_dh(gy.cuda(), sa.cuda(), w.cuda(), bias.cuda(), yg, _cu(mask), dhg, act)
    torch.testing.assert_close(dhg.cpu(), dh, rtol=1e-4, atol=1e-5)
    if cout <= 32:
        dsa, d_w, d_b = torch.zeros(rows, cin), torch.zeros(cin, cout), torch.zeros(cout)
        REF.sage_layer_bwd_narrow(gy, sa, w, bias, y, mask, dsa, d_w, d_b, act)
        outs = [torch.zeros_like(t).cuda() for t in (dsa, d_w, d_b)]
        cuda_backend.sage_layer_bwd_narrow(gy.cuda(), sa.cuda(), w.cuda(), bias.cuda(), yg, _cu(mask), *outs, act)
        for got, want in zip(outs, (dsa, d_w, d_b)):
            torch.testing.assert_close(got.cpu(), want, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize('act', [0, 1, 2])
def test_act_fwd_bwd(cuda_backend, act):
    g = torch.Generator().manual_seed(act)
    xw = torch.randn(211, 70, generator=g)
    x = xw[:, 3:64]                                              # strided view
    y = torch.zeros(211, 61)
    both(cuda_backend, 'act_fwd', [x, act, y], [2], rtol=1e-6, atol=1e-7)
    gy = torch.randn(211, 61, generator=g)
    gx = torch.zeros(211, 61)
    both(cuda_backend, 'act_bwd', [gy, y, act, gx], [3], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize('cin,cout', [(16, 20), (20, 20), (60, 20), (33, 32), (5, 3)])
@pytest.mark.parametrize('masked', [False, True])
def test_sage_layer_bwd_narrow(cuda_backend, cin, cout, masked):
    """Fused narrow backward (dh, dW, db, dsa in one pass) == the unfused formulas."""
    b, nmax = 4, 611
    rows = b * nmax
    g = torch.Generator().manual_seed(cin * 100 + cout)
    sa = torch.randn(rows, cin, generator=g)
    w = torch.randn(cin, cout, generator=g) / cin ** 0.5
    bias = torch.randn(cout, generator=g) * 0.1
    mask = _mask(b, nmax) if masked else None
    y = torch.zeros(rows, cout)
    REF.sage_epilogue_fwd(sa, w, bias, mask, y)
    gyw = torch.randn(rows, cout + 12, generator=g)
    gy = gyw[:, 4:4 + cout]                              # strided like a slice of the concat gradient
    both(cuda_backend, 'sage_layer_bwd_narrow',
         [gy, sa, w, bias, y, mask, torch.zeros(rows, cin), torch.zeros(cin, cout), torch.zeros(cout)],
         [6, 7, 8], rtol=2e-4, atol=2e-4)


def test_scale_rows_bwd(cuda_backend):
    g = torch.Generator().manual_seed(1)
    dsa, sa = torch.randn(500, 60, generator=g), torch.randn(500, 60, generator=g)
    rs = torch.rand(500, generator=g) + 0.1
    both(cuda_backend, 'scale_rows_bwd', [dsa, sa, rs, torch.zeros(500, 60), torch.randn(500, generator=g)],
         [3, 4], rtol=1e-4, atol=1e-5)


# ------------------------------------------------------------------------------- batch norm
@pytest.mark.parametrize('c', [20, 114, 300, 1140])            # 300, 1140: the float4 column reductions
@pytest.mark.parametrize('masked', [False, True])
def test_batchnorm(cuda_backend, c, masked):
    b, nmax = 4, 701
    rows = b * nmax
    g = torch.Generator().manual_seed(c)
    y = torch.rand(rows, c, generator=g)
    mask = _mask(b, nmax) if masked else None
    rmean, rvar = torch.randn(c, generator=g), torch.rand(c, generator=g) + 0.5
    mean, invstd = torch.zeros(c), torch.zeros(c)
    both(cuda_backend, 'bn_stats', [y, rmean, rvar, 0.1, 1e-5, mean, invstd], [1, 2, 5, 6], rtol=2e-5, atol=1e-6)
    gamma, beta = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g)
    wide = torch.zeros(rows, c + 40)
    both(cuda_backend, 'bn_apply', [y, mean, invstd, gamma, beta, mask, wide[:, 20:20 + c]], [6], rtol=1e-5, atol=1e-5)
    both(cuda_backend, 'bn_apply', [y, None, None, None, None, mask, torch.zeros(rows, c)], [6], rtol=0, atol=0)
    gx = torch.randn(rows, c + 40, generator=g)[:, 20:20 + c]
    for training in (True, False):
        both(cuda_backend, 'bn_bwd', [gx, y, mean, invstd, gamma, mask, training, torch.zeros(rows, c),
                                      torch.zeros(c), torch.zeros(c)], [7, 8, 9], rtol=2e-4, atol=2e-4)
    e_mean, e_inv = torch.zeros(c), torch.zeros(c)
    both(cuda_backend, 'bn_eval_stats', [rmean, rvar, 1e-5, e_mean, e_inv], [3, 4], rtol=1e-6, atol=0)
    both(cuda_backend, 'colsum', [gx, mask, torch.zeros(c)], [2], rtol=1e-4, atol=1e-4)


# ------------------------------------------------------------------------------- softmax / read-out
@pytest.mark.parametrize('c', [3, 114, 1140])
def test_softmax(cuda_backend, c):
    b, nmax = 3, 97
    rows = b * nmax
    g = torch.Generator().manual_seed(c)
    z = torch.randn(rows, c, generator=g) * 3
    mask = _mask(b, nmax)
    s, sf, arg = torch.zeros(rows, c), torch.zeros(rows, c), torch.zeros(rows, dtype=torch.int32)
    both(cuda_backend, 'softmax_fwd', [z, mask, s, sf, arg], [2, 3, 4], rtol=1e-5, atol=1e-7)   # arg-max bit-exact
    both(cuda_backend, 'softmax_fwd', [z, None, s, None, None], [2], rtol=1e-5, atol=1e-7)
    gs = torch.randn(rows, c, generator=g)
    both(cuda_backend, 'softmax_bwd', [gs, s, torch.zeros(rows, c)], [2], rtol=1e-4, atol=1e-6)
    # padded leading dimension (what ops._mat hands out for 114 clusters: ld 116)
    ld = (c + 3) // 4 * 4 + 4
    sp, gsp = torch.zeros(rows, ld)[:, :c], torch.randn(rows, ld, generator=g)[:, :c]
    both(cuda_backend, 'softmax_fwd', [z, mask, sp, None, None], [2], rtol=1e-5, atol=1e-7)
    both(cuda_backend, 'softmax_bwd', [gsp, sp, torch.zeros(rows, c)], [2], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize('c', [60, 20, 7])
def test_colmax(cuda_backend, c):
    b, n = 5, 333
    g = torch.Generator().manual_seed(c)
    x = torch.randn(b * n, c, generator=g)
    x[n:n + 40] = 0                                      # ties among zero rows: first index wins
    out, arg = torch.zeros(b, c), torch.zeros(b, c, dtype=torch.int32)
    both(cuda_backend, 'colmax_fwd', [x, b, n, out, arg], [3, 4], rtol=0, atol=0)
    gg = torch.randn(b, c, generator=g)
    both(cuda_backend, 'colmax_bwd', [gg, arg, b, n, torch.zeros(b * n, c)], [4], rtol=0, atol=0)


# ------------------------------------------------------------------------------- dense adjacency helpers
@pytest.mark.parametrize('n', [114, 1140, 37])
def test_dense_adjacency_helpers(cuda_backend, n):
    b = 3
    g = torch.Generator().manual_seed(n)
    adj = torch.rand(b, n, n, generator=g) * (4.0 / n)
    adj[0, 3] *= 0.01                                    # a row whose sum is < 1 (clamped)
    rs, rsum = torch.zeros(b * n), torch.zeros(b * n)
    both(cuda_backend, 'row_inv_clamped_sum', [adj, rs, rsum], [1, 2], rtol=1e-5, atol=1e-6)
    d_adj, d_rs = torch.randn(b, n, n, generator=g), torch.randn(b * n, generator=g)
    both(cuda_backend, 'add_rs_grad', [d_adj, rsum, d_rs], [0], rtol=1e-5, atol=1e-6)
    for mask in (None, _mask(b, n)):
        out, rho = torch.zeros(b, n, n), torch.zeros(b * n)
        both(cuda_backend, 'dense_renorm_fwd', [adj, 0.4, mask, out, rho], [3, 4], rtol=1e-5, atol=1e-7)
        gout = torch.randn(b, n, n, generator=g)
        both(cuda_backend, 'dense_renorm_bwd', [gout, out, rho, 0.4, mask, torch.zeros(b, n, n)], [5],
             rtol=1e-4, atol=1e-5)


# ------------------------------------------------------------------------------- dense contractions
GEMM_SHAPES = [  # (batch, m, n, k)
    (1, 200, 1140, 1180), (3, 114, 114, 300), (2, 1140, 60, 500), (1, 20, 1140, 5000), (2, 37, 154, 33),
    (1, 1000, 20, 1140), (4, 130, 130, 20), (1, 1180, 1140, 9000),
]


@pytest.mark.parametrize('shape', GEMM_SHAPES)
@pytest.mark.parametrize('ta,tb', [(False, False), (False, True), (True, False), (True, True)])
def test_gemm_fp32(cuda_backend, shape, ta, tb):
    """fp32 FMA path (precision 0), all four transposes, unaligned shapes, split-K."""
    batch, m, n, k = shape
    g = torch.Generator().manual_seed(m * 7 + n * 3 + k)
    a = torch.randn(batch, *((k, m) if ta else (m, k)), generator=g)
    b = torch.randn(batch, *((n, k) if tb else (k, n)), generator=g)
    c = torch.zeros(batch, m, n)
    if batch == 1:
        a, b, c = a[0], b[0], c[0]
    tol = 2e-5 * k ** 0.5
    both(cuda_backend, 'gemm', [a, b, c, ta, tb, 1.0, 0.0, None, None, None, 0], [2], rtol=1e-4, atol=tol)


def test_gemm_epilogue_options(cuda_backend):
    batch, m, n, k = 3, 53, 70, 41
    g = torch.Generator().manual_seed(3)
    a, b = torch.randn(batch, m, k, generator=g), torch.randn(batch, k, n, generator=g)
    c = torch.randn(batch, m, n, generator=g)
    bias, rsc = torch.randn(n, generator=g), torch.rand(batch * m, generator=g)
    mask = _mask(batch, m)
    both(cuda_backend, 'gemm', [a, b, c, False, False, 0.5, 1.0, bias, rsc, mask, 0], [2], rtol=1e-4, atol=1e-4)
    # strided operands / output: column slices of wider buffers
    aw, cw = torch.randn(batch * m, 100, generator=g), torch.zeros(batch * m, 90)
    a2 = aw[:, 8:8 + k].as_strided((batch, m, k), (m * 100, 100, 1), 8)
    c2 = cw[:, 4:4 + n].as_strided((batch, m, n), (m * 90, 90, 1), 4)
    both(cuda_backend, 'gemm', [a2, b, c2, False, False, 1.0, 0.0, None, None, None, 0], [2], rtol=1e-4, atol=1e-4)


# ------------------------------------------------------------------------------- tcgen05 3xTF32 GEMM
TC_SHAPES = [(1, 256, 256, 128), (1, 1000, 1140, 1180), (3, 1140, 1140, 1800), (1, 1140, 1180, 20000),
             (2, 200, 136, 72),
             # skinny outputs -> the BN=32 / BN=64 instantiations (dense-stage A'x, S gX', split-K variants)
             (3, 1140, 20, 1140), (2, 1800, 60, 1140), (1, 300, 32, 9000), (2, 132, 64, 520), (1, 1140, 36, 20000)]


@pytest.mark.parametrize('shape', TC_SHAPES)
@pytest.mark.parametrize('ta,tb', [(False, False), (False, True), (True, False), (True, True)])
def test_gemm_tcgen05(cuda_backend, shape, ta, tb):
    """precision=1 forces the tensor-core path (K-major and MN-major operands, K tails, split-K).
    3xTF32 restores the fp32 products; what remains is the RZ accumulation of the tensor core (see assert)."""
    batch, m, n, k = shape
    g = torch.Generator().manual_seed(m + n + k)
    a = torch.randn(batch, *((k, m) if ta else (m, k)), generator=g)
    b = torch.randn(batch, *((n, k) if tb else (k, n)), generator=g)
    want = ((a.transpose(1, 2) if ta else a).double() @ (b.transpose(1, 2) if tb else b).double())
    c = torch.full((batch, m, n), float('nan')).cuda()
    ag, bg = a.cuda(), b.cuda()
    if batch == 1:
        ag, bg, c, want = ag[0], bg[0], c[0], want[0]
    cuda_backend.gemm(ag, bg, c, ta, tb, precision=1)
    torch.cuda.synchronize()
    err = float((c.cpu().double() - want).abs().max())
    # the tensor core accumulates into TMEM with round-toward-zero, so the kernel promotes the accumulator
    # to fp32 registers every K = 128; what is left is fp32-grade: < 5e-6 of the output scale
    # (un-promoted 3xTF32 measured ~1.3e-5, single-pass TF32 ~1e-3)
    assert err <= 5e-6 * float(want.abs().max()), 'max abs err %.3e vs scale %.3e (k=%d)' % (
        err, float(want.abs().max()), k)


def test_gemm_tcgen05_epilogue(cuda_backend):
    """alpha / beta / bias / row_scale / row mask and strided operands on the tensor-core path."""
    batch, m, n, k = 2, 300, 260, 200
    g = torch.Generator().manual_seed(9)
    aw = torch.randn(batch * m, k + 24, generator=g)
    a = aw[:, 8:8 + k].as_strided((batch, m, k), (m * (k + 24), k + 24, 1), 8)
    b = torch.randn(batch, k, n, generator=g)
    c0 = torch.randn(batch, m, n, generator=g)
    bias, rsc = torch.randn(n, generator=g), torch.rand(batch * m, generator=g)
    mask = _mask(batch, m)
    want = c0.clone()
    REF.gemm(a, b, want, False, False, 0.5, 1.0, bias, rsc, mask)
    cg = c0.cuda()
    awg = aw.cuda()
    ag = awg.as_strided((batch, m, k), (m * (k + 24), k + 24, 1), 8)
    cuda_backend.gemm(ag, b.cuda(), cg, False, False, 0.5, 1.0, bias.cuda(), rsc.cuda(), _cu(mask), precision=1)
    torch.cuda.synchronize()
    torch.testing.assert_close(cg.cpu(), want, rtol=1e-4, atol=2e-4)


def test_tf32_operand_truncation_probe(cuda_backend):
    """Records how kind::tf32 reads fp32 operands (the split writes an explicit hi so either is handled)."""
    m = n = 128
    k = 64
    a = torch.zeros(m, k)
    a[:, 0] = 1.0 + 2.0 ** -11 + 2.0 ** -12          # low bits beyond the 10-bit TF32 mantissa
    b = torch.zeros(k, n)
    b[0, :] = 1.0
    c = torch.zeros(m, n).cuda()
    cuda_backend.gemm(a.cuda(), b.cuda(), c, False, False, precision=1)
    torch.cuda.synchronize()
    assert float((c.cpu() - a[0, 0]).abs().max()) < 1e-7        # 3xTF32 reproduces all fp32 bits here


# ------------------------------------------------------------------------------- DenseJK (fused bi-LSTM attention)
@pytest.mark.parametrize('rows,c', [(1, 20), (200, 20), (1000, 20), (333, 16), (333, 32)])
def test_jk_lstm_fwd_bwd(cuda_backend, rows, c):
    """cgc_jk_lstm_fwd / bwd against the CPU formulas (themselves pinned to torch.nn.LSTM autograd, see
    test_oracle.py::test_jk_reference_formulas_match_autograd); strided feats (a slice of a wider buffer); the
    channel widths the fused kernels are instantiated for (--hidden-dim 16 / 20 / 32)."""
    layers = 3
    hid = c * layers // 2
    assert cuda_backend.jk_supported(c, layers) and not cuda_backend.jk_supported(64, 3)
    assert not cuda_backend.jk_supported(20, 2)
    torch.manual_seed(rows)
    lstm = torch.nn.LSTM(c, hid, bidirectional=True, batch_first=True)
    att = torch.nn.Linear(2 * hid, 1)
    names = ['weight_ih_l0', 'weight_hh_l0', 'bias_ih_l0', 'bias_hh_l0']
    params = [getattr(lstm, n).detach() for n in names] + [getattr(lstm, n + '_reverse').detach() for n in names]
    wide = torch.randn(rows, layers * c + 8)
    feats = wide[:, 4:4 + layers * c]
    out, tape = torch.zeros(rows, c), torch.zeros(REF.jk_tape_floats(rows, c, layers))
    assert cuda_backend.jk_tape_floats(rows, c, layers) == tape.numel()
    REF.jk_lstm_fwd(feats, c, layers, params, att.weight.detach(), att.bias.detach(), out, tape)
    fg = wide.cuda()[:, 4:4 + layers * c]
    pg = [p.cuda() for p in params]
    outg, tapeg = torch.zeros(rows, c).cuda(), torch.zeros(tape.numel()).cuda()
    cuda_backend.jk_lstm_fwd(fg, c, layers, pg, att.weight.detach().cuda(), att.bias.detach().cuda(), outg, tapeg)
    torch.testing.assert_close(outg.cpu(), out, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(tapeg.cpu(), tape, rtol=1e-5, atol=2e-6)
    # inference call (no tape) gives the same output
    out2 = torch.zeros(rows, c).cuda()
    cuda_backend.jk_lstm_fwd(fg, c, layers, pg, att.weight.detach().cuda(), att.bias.detach().cuda(), out2, None)
    assert torch.equal(out2, outg)
    g_out = torch.randn(rows, c)
    gf, dg, dsc = torch.zeros(rows, layers * c), torch.zeros(2, layers, 4 * hid, rows), torch.zeros(layers, rows)
    REF.jk_lstm_bwd(g_out, feats, c, layers, params, att.weight.detach(), tape, gf, dg, dsc)
    gfg, dgg, dscg = [torch.zeros_like(t).cuda() for t in (gf, dg, dsc)]
    cuda_backend.jk_lstm_bwd(g_out.cuda(), fg, c, layers, pg, att.weight.detach().cuda(), tape.cuda(), gfg, dgg, dscg)
    torch.testing.assert_close(gfg.cpu(), gf, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(dgg.cpu(), dg, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(dscg.cpu(), dsc, rtol=1e-4, atol=1e-6)


# ------------------------------------------------------------------------------- radius_graph (loader side)
@pytest.mark.parametrize('sizes,r,k,loop', [([1800, 1800, 700], 100.0, 8, True), ([300, 41, 1], 30.0, 8, True),
                                            ([500, 260], 60.0, 12, False), ([2999], 45.0, 20, True)])
def test_radius_graph_matches_ckdtree(cuda_backend, sizes, r, k, loop):
    """GPU radius_graph == scipy cKDTree.query (the CPU path of torch_cluster 1.4.2), bit-exact neighbour lists
    (nearest first), then the COO edge_index the reference's loaders emit (dataflow/data.py:348)."""
    g = torch.Generator().manual_seed(len(sizes) * 1000 + k)
    pos = torch.rand(sum(sizes), 2, generator=g) * 1000.0
    start = torch.zeros(len(sizes) + 1, dtype=torch.int32)
    start[1:] = torch.tensor(sizes).cumsum(0).to(torch.int32)
    nbr = torch.zeros(sum(sizes), k, dtype=torch.int32)
    deg = torch.zeros(sum(sizes), dtype=torch.int32)
    both(cuda_backend, 'radius_graph', [pos, start, len(sizes), max(sizes), r, k, loop, nbr, deg], [7, 8])
    assert int(deg.max()) <= k and (loop is False or int(deg.min()) >= 1)
    ei = graph.radius_graph(pos.cuda(), r, None, loop, k, num_nodes_list=sizes).cpu()
    valid = nbr >= 0
    centre = torch.arange(sum(sizes)).view(-1, 1).expand_as(nbr)
    assert torch.equal(ei[0], nbr[valid].long()) and torch.equal(ei[1], centre[valid])
    if loop:
        assert torch.equal(nbr[:, 0].long(), torch.arange(sum(sizes)))      # the centre itself comes first


@pytest.mark.parametrize('sizes,ratio', [([1800, 700, 33], 0.5), ([2999], 0.35), ([64, 1, 5], 1.0)])
def test_farthest_point_sample_matches_reference_sampler(cuda_backend, sizes, ratio):
    """cgc_farthest_point_sample == FarthestSampler on the int16 distance matrix (bit-exact index sequences)."""
    g = torch.Generator().manual_seed(sum(sizes))
    pos = torch.rand(sum(sizes), 2, generator=g) * 3584.0
    start = torch.zeros(len(sizes) + 1, dtype=torch.int32)
    start[1:] = torch.tensor(sizes).cumsum(0).to(torch.int32)
    ks = [max(1, int(n * ratio)) for n in sizes]
    ostart = torch.zeros(len(sizes) + 1, dtype=torch.int32)
    ostart[1:] = torch.tensor(ks).cumsum(0).to(torch.int32)
    first = torch.tensor([int(torch.randint(0, n, (1,), generator=g)) for n in sizes], dtype=torch.int32)
    out = torch.zeros(sum(ks), dtype=torch.int32)
    both(cuda_backend, 'farthest_point_sample', [pos, start, first, ostart, max(sizes), out], [5])
    for gi, k in enumerate(ks):
        ids = out[int(ostart[gi]):int(ostart[gi + 1])]
        assert int(ids[0]) == int(first[gi]) and int(ids.max()) < sizes[gi]
    picks = graph.sample_nodes(pos.cuda(), sizes, 0.5, 'fuse', generator=torch.Generator().manual_seed(1))
    for n, ids in zip(sizes, picks):
        assert ids.numel() == int(n * 0.5) and ids.unique().numel() == ids.numel() and (n == 1 or int(ids.max()) < n)


@pytest.mark.parametrize('cin,cout', [(16, 20), (20, 20), (60, 20), (33, 32), (5, 3)])
@pytest.mark.parametrize('act', [0, 1])
def test_sage_layer_fwd_narrow(cuda_backend, cin, cout, act):
    """Fused narrow forward (epilogue + BatchNorm statistics in one pass) == epilogue followed by bn_stats;
    strided sa / y (column slices), ragged row mask, several tiles per CTA."""
    b, nmax = 5, 1237
    rows = b * nmax
    g = torch.Generator().manual_seed(cin * 100 + cout)
    saw = torch.randn(rows, cin + 8, generator=g)
    sa = saw[:, 4:4 + cin]
    w = torch.randn(cin, cout, generator=g) / cin ** 0.5
    bias = torch.randn(cout, generator=g) * 0.1
    mask = _mask(b, nmax)
    yw = torch.zeros(rows, cout + 4)
    y = yw[:, 0:cout]
    rmean, rvar = torch.randn(cout, generator=g), torch.rand(cout, generator=g) + 0.5
    mean, invstd = torch.zeros(cout), torch.zeros(cout)
    both(cuda_backend, 'sage_layer_fwd_narrow', [sa, w, bias, mask, y, act, rmean, rvar, 0.1, 1e-5, mean, invstd],
         [4, 6, 7, 10, 11], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('feat,b,nmax,width,maxdeg', [(1140, 3, 200, 8, 8), (116, 5, 333, 16, 18), (64, 2, 1800, 8, 10),
                                                      (300, 1, 3500, 8, 8), (1140, 8, 1800, 8, 8)])
@pytest.mark.parametrize('weighted,accumulate', [(False, False), (True, True)])
def test_gather_wide_ell(cuda_backend, feat, b, nmax, width, maxdeg, weighted, accumulate):
    """Shared-memory wide gather (2-D TMA slices, double-buffered) == the CPU formulas == the generic CSR gather;
    strided x / out (leading dimension padded), overflow rows, += mode, single-buffer fallback (nmax = 3500)."""
    rows = b * nmax
    rowptr, col = _random_csr(rows, maxdeg, seed=nmax + feat, block=nmax) if rows < 5000 else _knn_csr(b, nmax)
    nnz = int(rowptr[-1])
    g = torch.Generator().manual_seed(9)
    ld = feat + 4
    xw = torch.randn(rows, ld, generator=g)
    x = xw[:, :feat]
    val = torch.rand(nnz + 3, generator=g) if weighted else None
    trows = REF.ell16_rows(b, nmax)
    ell = torch.zeros(trows * width, dtype=torch.int16)
    ellw = torch.zeros(trows * width) if weighted else None
    ells = torch.zeros(trows)
    perm = torch.zeros(trows // 8, dtype=torch.int32)
    REF.csr_to_ell16(rowptr, col, val, None, b, nmax, width, 1, ell, ellw, ells, perm)
    scale = ells if weighted else None
    ow = torch.randn(rows, ld, generator=g)
    out = ow[:, :feat]
    args = both(cuda_backend, 'gather_wide_ell',
                [ell, ellw, scale, perm, width, rowptr, col, val, x, b, nmax, out, accumulate], [11],
                rtol=1e-5, atol=1e-5)
    assert torch.equal(args[11]._base[:, feat:].cpu(), ow[:, feat:])            # nothing outside the slice touched


def _ell_tables(rowptr, col, val, rsc, b, nmax, width, pair):
    trows = REF.ell16_rows(b, nmax)
    ell = torch.zeros(trows * width, dtype=torch.int16)
    ellw = torch.zeros(trows * width) if val is not None else None
    ells = torch.zeros(trows)
    perm = torch.zeros(trows // 8, dtype=torch.int32)
    REF.csr_to_ell16(rowptr, col, val, rsc, b, nmax, width, pair, ell, ellw, ells, perm)
    return ell, ellw, ells, perm


@pytest.mark.parametrize('feat,b,nmax,width,maxdeg', [(16, 3, 70, 8, 8), (20, 5, 333, 8, 10), (16, 64, 1800, 8, 8),
                                                      (20, 64, 1800, 8, 8), (8, 9, 40, 16, 16), (32, 2, 1500, 16, 18)])
@pytest.mark.parametrize('weighted,accumulate,ready', [(False, False, True), (True, True, False), (False, True, True)])
def test_sage_gather_ell_strided(cuda_backend, feat, b, nmax, width, maxdeg, weighted, accumulate, ready):
    """General ELL16 gather: x / out as column slices of wider matrices (2-D TMA staging), += mode, and the
    programmatic-launch prologue with and without early table loads; nothing outside the slices is touched."""
    rows = b * nmax
    rowptr, col = _random_csr(rows, maxdeg, seed=nmax + feat, block=nmax) if rows < 5000 else _knn_csr(b, nmax)
    nnz = int(rowptr[-1])
    g = torch.Generator().manual_seed(21)
    val = torch.rand(nnz + 3, generator=g) if weighted else None
    rsc = torch.rand(rows, generator=g) if weighted else None
    ell, ellw, ells, perm = _ell_tables(rowptr, col, val, rsc, b, nmax, width, int(feat == 16))
    xw = torch.randn(rows, feat + 44, generator=g)
    x = xw[:, 20:20 + feat]
    ow = torch.randn(rows, feat + 12, generator=g)
    out = ow[:, 4:4 + feat]
    want = out.clone() if accumulate else torch.zeros(rows, feat)
    ref = torch.zeros(rows, feat)
    REF.sage_gather_mean_fwd(rowptr, col, val, rsc, x.contiguous(), b, nmax, ref)
    want = want + ref if accumulate else ref
    args = both(cuda_backend, 'sage_gather_ell',
                [ell, ellw, ells, perm, width, rowptr, col, val, x, b, nmax, out, accumulate, ready], [11],
                rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(out, want, rtol=1e-5, atol=1e-5)
    base = args[11]._base.cpu()
    assert torch.equal(base[:, :4], ow[:, :4]) and torch.equal(base[:, 4 + feat:], ow[:, 4 + feat:])
    # contiguous operands through the same entry point (1-D bulk-copy staging)
    xc, oc = x.contiguous(), (out.clone() if accumulate else torch.zeros(rows, feat))
    both(cuda_backend, 'sage_gather_ell', [ell, ellw, ells, perm, width, rowptr, col, val, xc, b, nmax, oc,
                                           accumulate, ready], [11], rtol=1e-5, atol=1e-5)
    if not weighted:
        # unweighted table without a scale vector: the DenseSAGEConv mean 1 / max(deg, 1) from the row's own slots
        # (feat 16, width 8, contiguous, plain store = the two-CTAs-per-SM roofline variant)
        ell, ellw, _, perm = _ell_tables(rowptr, col, None, None, b, nmax, width, int(feat == 16))
        for xa, oa in ((xc, torch.zeros(rows, feat)), (x, ow.clone()[:, 4:4 + feat])):
            if accumulate:
                oa.zero_()
            both(cuda_backend, 'sage_gather_ell', [ell, None, None, perm, width, rowptr, col, None, xa, b, nmax, oa,
                                                   False, ready], [11], rtol=1e-5, atol=1e-5)
            deg = (rowptr[1:] - rowptr[:-1]).float().clamp(min=1)
            mean = torch.zeros(rows, feat)
            REF.spmm(rowptr, col, None, xc, None, 1.0 / deg, mean, False)
            torch.testing.assert_close(oa, mean, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('cin,b,nmax,width,maxdeg', [(16, 3, 70, 8, 8), (20, 5, 333, 8, 10), (16, 64, 1800, 8, 8),
                                                     (20, 64, 1800, 8, 8), (16, 7, 1237, 16, 18), (20, 2, 2100, 16, 12)])
@pytest.mark.parametrize('weighted,strided,stats,act', [(False, False, True, 0), (True, True, True, 1),
                                                        (False, True, False, 2)])
def test_sage_layer_ell_fwd(cuda_backend, cin, b, nmax, width, maxdeg, weighted, strided, stats, act):
    """Fused narrow layer (gather + W + bias + L2 norm + mask + activation + BatchNorm statistics) == ELL16 gather
    followed by the epilogue and bn_stats: aggregate sa, output y (strided), mean / invstd, running statistics;
    ragged row mask, overflow rows (width 16 with degree up to 18), both tile-staging paths, eval mode (no stats)."""
    rows = b * nmax
    rowptr, col = _random_csr(rows, maxdeg, seed=nmax + cin, block=nmax) if rows < 5000 else _knn_csr(b, nmax)
    nnz = int(rowptr[-1])
    g = torch.Generator().manual_seed(cin + nmax)
    val = torch.rand(nnz + 3, generator=g) if weighted else None
    rsc = torch.rand(rows, generator=g) if weighted else None
    ell, ellw, ells, perm = _ell_tables(rowptr, col, val, rsc, b, nmax, width, int(cin == 16))
    xw = torch.randn(rows, cin + 40, generator=g)
    x = xw[:, 20:20 + cin] if strided else xw[:, :cin].contiguous()
    w = torch.randn(cin, 20, generator=g) / cin ** 0.5
    bias = torch.randn(20, generator=g) * 0.1
    mask = _mask(b, nmax, seed=3)
    sa = torch.zeros(rows, cin)
    yw = torch.zeros(rows, 24)
    y = yw[:, :20]
    rmean, rvar = torch.randn(20, generator=g), torch.rand(20, generator=g) + 0.5
    mean, invstd = (torch.zeros(20), torch.zeros(20)) if stats else (None, None)
    outs = [15, 16] + ([17, 18, 21, 22] if stats else [])
    both(cuda_backend, 'sage_layer_ell_fwd',
         [ell, ellw, ells, perm, width, rowptr, col, val, x, b, nmax, w, bias, mask, act, sa, y,
          rmean if stats else None, rvar if stats else None, 0.1, 1e-5, mean, invstd, True], outs,
         rtol=2e-5, atol=2e-6)
    # no bias, tables flagged as fresh (no early loads); unweighted tables: no scale vector (mean from the slots)
    y2 = torch.zeros(rows, 20)
    if not weighted:
        ell, ellw, ells, perm = _ell_tables(rowptr, col, None, None, b, nmax, width, int(cin == 16))
        ells = None
    both(cuda_backend, 'sage_layer_ell_fwd',
         [ell, ellw, ells, perm, width, rowptr, col, val, x, b, nmax, w, None, mask, act, sa, y2,
          None, None, 0.1, 1e-5, None, None, False], [15, 16], rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize('case', ['lin', 'splitk', 'batched_pair', 'single_pass'])
def test_gemm_tcgen05_bitwise_determinism(cuda_backend, case):
    """The CTA-pair kernel hands tensor / shared memory between the two CTAs with CTA-scope arrives ordered by
    tcgen05 fences.  A race there would show up as run-to-run differences: 100 repeats of
    the same contraction must be BITWISE identical -- persistent tile loop (lin), split-K with its fixed-order
    reduction, the batched pair layout of the pooled adjacency, and the single-pass mode."""
    g = torch.Generator().manual_seed(3)
    prec = 1
    if case == 'lin':
        a, b, c = torch.randn(9000, 1180, generator=g), torch.randn(1140, 1180, generator=g), torch.empty(9000, 1140)
        ta, tb = False, True
    elif case == 'splitk':                     # dW of lin: K = rows, few output tiles -> split-K partials
        a, b, c = torch.randn(20000, 1140, generator=g), torch.randn(20000, 1180, generator=g), torch.empty(1140, 1180)
        ta, tb = True, False
    elif case == 'batched_pair':               # S^T (A S): batch of [1140, 1800] x [1800, 1140]
        a, b = torch.rand(3, 1800, 1140, generator=g), torch.rand(3, 1800, 1140, generator=g)
        c, ta, tb = torch.empty(3, 1140, 1140), True, False
    else:
        a, b, c = torch.randn(9000, 1140, generator=g), torch.randn(1140, 1180, generator=g), torch.empty(9000, 1180)
        ta, tb, prec = False, False, 2
    a, b, c = a.cuda(), b.cuda(), c.cuda()
    cuda_backend.gemm(a, b, c, ta, tb, precision=prec)
    first = c.clone()
    for _ in range(100):
        c.zero_()
        cuda_backend.gemm(a, b, c, ta, tb, precision=prec)
        assert torch.equal(c, first), 'tcgen05 GEMM (%s) is not run-to-run deterministic' % case
    ref = (a.transpose(-1, -2) if ta else a).double() @ (b.transpose(-1, -2) if tb else b).double()
    tol = 2e-3 if prec == 2 else 5e-6
    assert float((first.double() - ref).abs().max() / ref.abs().max()) < tol


def test_gemm_single_pass_tf32(cuda_backend):
    """precision = 2: one TF32 pass on round-to-nearest operands.  Unbiased (sign-definite operands show no drift:
    relative error of a sum of positive products << 2^-11) and ~1e-4-grade on random data, all operand layouts."""
    g = torch.Generator().manual_seed(5)
    for ta in (False, True):
        for tb in (False, True):
            m, n, k = 700, 1140, 1800
            a = torch.rand(k, m, generator=g) if ta else torch.rand(m, k, generator=g)
            b = torch.rand(n, k, generator=g) if tb else torch.rand(k, n, generator=g)
            c = torch.empty(m, n).cuda()
            cuda_backend.gemm(a.cuda(), b.cuda(), c, ta, tb, precision=2)
            ref = (a.t() if ta else a).double() @ (b.t() if tb else b).double()
            rel = (c.cpu().double() - ref) / ref
            assert float(rel.abs().max()) < 1e-4, (ta, tb, float(rel.abs().max()))      # K = 1800 averages the rounding
            assert abs(float(rel.mean())) < 3e-6, (ta, tb, float(rel.mean()))            # no truncation bias (2^-12 = 2.4e-4)


@pytest.mark.parametrize('cin,cout', [(20, 1140), (20, 114), (60, 300), (16, 33)])
@pytest.mark.parametrize('act', [0, 1])
def test_sage_layer_fwd_wide(cuda_backend, cin, cout, act):
    """Wide epilogue with the BatchNorm statistics accumulated in the column-owning threads == epilogue followed by
    bn_stats (many tiles per CTA, ragged mask, strided y)."""
    b, nmax = 7, 1237
    rows = b * nmax
    g = torch.Generator().manual_seed(cin * 100 + cout)
    sa = torch.randn(rows, cin, generator=g)
    w = torch.randn(cin, cout, generator=g) / cin ** 0.5
    bias = torch.randn(cout, generator=g) * 0.1
    mask = _mask(b, nmax)
    yw = torch.zeros(rows, cout + 4)
    y = yw[:, 0:cout]
    rmean, rvar = torch.randn(cout, generator=g), torch.rand(cout, generator=g) + 0.5
    mean, invstd = torch.zeros(cout), torch.zeros(cout)
    both(cuda_backend, 'sage_layer_fwd_wide', [sa, w, bias, mask, y, act, rmean, rvar, 0.1, 1e-5, mean, invstd],
         [4, 6, 7, 10, 11], rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize('cin,cout', [(20, 1140), (20, 114), (60, 300)])
@pytest.mark.parametrize('training,act', [(True, 0), (False, 1)])
def test_sage_dh_bn(cuda_backend, cin, cout, training, act):
    """Epilogue backward with the BatchNorm backward folded in == bn_bwd (apply step) followed by sage_dh;
    bn_bwd with gy = None produces the same column sums (ragged mask, strided gradient slice)."""
    b, nmax = 5, 653
    rows = b * nmax
    g = torch.Generator().manual_seed(cin + cout)
    sa = torch.randn(rows, cin, generator=g)
    w = torch.randn(cin, cout, generator=g) / cin ** 0.5
    bias = torch.randn(cout, generator=g) * 0.1
    mask = _mask(b, nmax, seed=2)
    y = torch.zeros(rows, cout)
    REF.sage_epilogue_fwd(sa, w, bias, mask, y, act)
    mean, var = y.mean(0), y.var(0, unbiased=False)
    invstd = 1.0 / torch.sqrt(var + 1e-5)
    gamma = torch.rand(cout, generator=g) + 0.5
    gxw = torch.randn(rows, cout + 40, generator=g)
    gx = gxw[:, 40:40 + cout]
    dg, db = torch.zeros(cout), torch.zeros(cout)
    both(cuda_backend, 'bn_bwd', [gx, y, mean, invstd, gamma, mask, training, None, dg, db], [8, 9], rtol=1e-4, atol=1e-4)
    dh = torch.zeros(rows, cout)
    both(cuda_backend, 'sage_dh_bn', [gx, sa, w, bias, y, mean, invstd, gamma, dg, db, training, mask, dh, act], [12],
         rtol=1e-4, atol=2e-6)
    gy, dh2 = torch.zeros(rows, cout), torch.zeros(rows, cout)
    REF.bn_bwd(gx, y, mean, invstd, gamma, mask, training, gy, torch.zeros(cout), torch.zeros(cout))
    REF.sage_dh(gy, sa, w, bias, y, mask, dh2, act)
    torch.testing.assert_close(dh, dh2, rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize('cin,cout', [(20, 114), (16, 1138), (33, 301), (20, 1140)])
@pytest.mark.parametrize('fold', [False, True])
def test_sage_dh_row_pair_kernel_padded_width(cuda_backend, cin, cout, fold):
    """The warp-per-row-pair epilogue backward on column SLICES of wider 16-byte aligned buffers, widths that are not
    multiples of 4 included (stage 2 of config #3: 114): the columns next to the slice hold NaN (y, gy) or a sentinel
    (dh) -- nothing past the slice may be read into the result or written."""
    b, nmax = 3, 211
    rows = b * nmax
    g = torch.Generator().manual_seed(cin * cout)
    sa = torch.randn(rows, cin, generator=g)
    w = torch.randn(cin, cout, generator=g) / cin ** 0.5
    bias = torch.randn(cout, generator=g) * 0.1
    mask = _mask(b, nmax, seed=5)
    wide = ((cout + 3) // 4) * 4 + 8
    y_buf, g_buf, d_buf = (torch.full((rows, wide), float('nan')) for _ in range(3))
    d_buf.fill_(777.0)
    y, gy, dh = y_buf[:, 4:4 + cout], g_buf[:, 4:4 + cout], d_buf[:, 4:4 + cout]
    y_dense = torch.zeros(rows, cout)
    REF.sage_epilogue_fwd(sa, w, bias, mask, y_dense, 0)
    y.copy_(y_dense)
    gy.copy_(torch.randn(rows, cout, generator=g))
    if fold:
        mean, var = y_dense.mean(0), y_dense.var(0, unbiased=False)
        invstd = 1.0 / torch.sqrt(var + 1e-5)
        gamma = torch.rand(cout, generator=g) + 0.5
        dg, db = torch.randn(cout, generator=g), torch.randn(cout, generator=g)
        gpu = both(cuda_backend, 'sage_dh_bn', [gy, sa, w, bias, y, mean, invstd, gamma, dg, db, True, mask, dh, 0], [12],
                   rtol=1e-4, atol=2e-6)
        out = gpu[12]
    else:
        gpu = both(cuda_backend, 'sage_dh', [gy, sa, w, bias, y, mask, dh, 0], [6], rtol=1e-4, atol=2e-6)
        out = gpu[6]
    full = out._base if out._base is not None else out
    full = full.cpu().view(rows, wide)
    assert torch.isfinite(full[:, 4:4 + cout]).all()
    assert (full[:, :4] == 777.0).all() and (full[:, 4 + cout:] == 777.0).all(), 'wrote outside the slice'


@pytest.mark.parametrize('b,nmax,maxdeg', [(3, 70, 8), (64, 1800, 8), (5, 333, 10), (150, 40, 8), (2, 1750, 8)])
@pytest.mark.parametrize('ready', [True, False])
def test_sage_gather_ell_two_per_sm(cuda_backend, b, nmax, maxdeg, ready):
    """The two-CTAs-per-SM variant of the roofline configuration (pitch 8, unweighted, no scale vector, 16 contiguous
    features; with 2 or 4 CTAs per graph they form a cluster and share one multicast copy of the tile: 64 x 1800 -> 2,
    5 x 333 and 2 x 1750 -> 4) and the balanced row split (some graphs get one more CTA so that the grid covers every SM; 64 graphs:
    20 x 3 + 44 x 2 = 148) == the generic CSR mean; rows above the pitch walk the CSR (maxdeg 10)."""
    rows, feat = b * nmax, 16
    rowptr, col = _random_csr(rows, maxdeg, seed=nmax, block=nmax) if rows < 5000 else _knn_csr(b, nmax)
    ell, _, _, perm = _ell_tables(rowptr, col, None, None, b, nmax, 8, 1)
    x = torch.randn(rows, feat, generator=torch.Generator().manual_seed(3))
    deg = (rowptr[1:] - rowptr[:-1]).float().clamp(min=1)
    want = torch.zeros(rows, feat)
    REF.spmm(rowptr, col, None, x, None, 1.0 / deg, want, False)
    outs = []
    for two, pdl, mc in ((True, True, True), (True, True, False), (True, False, True), (False, True, True)):
        out = torch.zeros(rows, feat)
        gpu = both(cuda_backend, 'sage_gather_ell', [ell, None, None, perm, 8, rowptr, col, None, x, b, nmax, out, False,
                                                     ready, pdl, two, mc], [11], rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(out, want, rtol=1e-5, atol=1e-5)
        outs.append(gpu[11].cpu())
    assert all(torch.equal(outs[0], o) for o in outs[1:])        # same slot order in every kernel: bitwise equal
