"""Small invocations of the kernels that synchronise through mbarriers / TMA / tensor memory, for
compute-sanitizer (memcheck, racecheck, synccheck): the tcgen05 GEMM (single CTA, CTA pair, split-K, single pass,
all operand layouts), the ELL16 gathers (1-D bulk and 2-D TMA staging, programmatic launch), the fused layer and
the wide gather.  Results are checked against torch so a sanitizer-clean run is also a correct one."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cgc_net_b200 import _backend, graph, ops  # noqa: E402

be = _backend.get()
dev = 'cuda'
g = torch.Generator().manual_seed(0)


def check(name, got, want, tol):
    err = float((got.double().cpu() - want.double()).abs().max() / want.abs().max())
    print('%-28s rel err %.2e' % (name, err))
    assert err < tol, name


# tcgen05 GEMM
for ta in (False, True):
    for tb in (False, True):
        for prec in (1, 3):
            m, n, k = 384, 256, 160
            a = torch.randn((k, m) if ta else (m, k), generator=g)
            b = torch.randn((n, k) if tb else (k, n), generator=g)
            c = torch.empty(m, n, device=dev)
            be.gemm(a.to(dev), b.to(dev), c, ta, tb, precision=prec)
            torch.cuda.synchronize()
            ref = (a.t() if ta else a).double() @ (b.t() if tb else b).double()
            check('gemm ta=%d tb=%d prec=%d' % (ta, tb, prec), c, ref, 1e-5 if prec == 1 else 5e-3)
a, b = torch.randn(8192, 200, generator=g), torch.randn(8192, 136, generator=g)       # split-K, pair
c = torch.empty(200, 136, device=dev)
be.gemm(a.to(dev), b.to(dev), c, True, False, precision=1)
torch.cuda.synchronize()
check('gemm split-K', c, a.double().t() @ b.double(), 1e-5)
a, b = torch.randn(3, 300, 96, generator=g), torch.randn(3, 96, 40, generator=g)      # batched, narrow tile
c = torch.empty(3, 300, 40, device=dev)
be.gemm(a.to(dev), b.to(dev), c, False, False, precision=1)
torch.cuda.synchronize()
check('gemm batched bn=64', c, a.double() @ b.double(), 1e-5)

# ELL16 gathers + fused layer + wide gather on a small ragged batch
graphs = [graph.synthetic_cell_graph(n, k, 16, seed=i, row_is_centre=(i % 2 == 0))
          for i, (n, k) in enumerate([(300, 8), (170, 5), (256, 12)])]
batch = graph.Batch.from_data_list(graphs).to(dev)
for norm_adj in (False, True):
    xp, sg = ops.build_sparse_graph(batch.x, batch.edge_index, batch.batch, batch.num_nodes_list, norm_adj)
    ref = torch.empty_like(xp)
    be.spmm(sg.rowptr, sg.col, sg.val, xp, None, sg.rs, ref, False)
    out = torch.empty_like(xp)
    be.sage_gather_ell(sg.ell, sg.ell_val, sg.ell_scale, sg.ell_perm, sg.ell_width, sg.rowptr, sg.col, sg.val, xp,
                       sg.num_graphs, sg.nmax, out, tables_ready=False)
    wide = torch.randn(xp.shape[0], 48, device=dev)
    wide[:, 8:24] = xp
    out2 = torch.zeros(xp.shape[0], 32, device=dev)
    be.sage_gather_ell(sg.ell, sg.ell_val, sg.ell_scale, sg.ell_perm, sg.ell_width, sg.rowptr, sg.col, sg.val,
                       wide[:, 8:24], sg.num_graphs, sg.nmax, out2[:, 4:20], tables_ready=True)
    torch.cuda.synchronize()
    check('gather ell (bulk) na=%d' % norm_adj, out, ref.cpu(), 1e-5)
    check('gather ell (2-D TMA) na=%d' % norm_adj, out2[:, 4:20], ref.cpu(), 1e-5)
    w, bias = torch.randn(16, 20, device=dev) * 0.3, torch.randn(20, device=dev) * 0.1
    sa, y = torch.empty_like(xp), torch.empty(xp.shape[0], 20, device=dev)
    rm, rv, mean, invstd = (torch.zeros(20, device=dev) for _ in range(4))
    be.sage_layer_ell_fwd(sg.ell, sg.ell_val, sg.ell_scale, sg.ell_perm, sg.ell_width, sg.rowptr, sg.col, sg.val, xp,
                          sg.num_graphs, sg.nmax, w, bias, sg.mask, 0, sa, y, rm, rv, 0.1, 1e-5, mean, invstd,
                          tables_ready=True)
    yref = torch.empty_like(y)
    be.sage_epilogue_fwd(ref, w, bias, sg.mask, yref, 0)
    torch.cuda.synchronize()
    check('fused layer na=%d' % norm_adj, y, yref.cpu(), 1e-4)
    s = torch.rand(xp.shape[0], 120, device=dev)[:, :116]
    p1, p2 = torch.empty(xp.shape[0], 120, device=dev)[:, :116], torch.empty(xp.shape[0], 120, device=dev)[:, :116]
    be.gather_wide_ell(sg.ell, sg.ell_val, None, sg.ell_perm, sg.ell_width, sg.rowptr, sg.col, sg.val, s,
                       sg.num_graphs, sg.nmax, p1, False)
    be.spmm(sg.rowptr, sg.col, sg.val, s, None, None, p2, False)
    torch.cuda.synchronize()
    check('wide gather na=%d' % norm_adj, p1, p2.cpu(), 1e-5)
# round-2b kernels: two-per-SM gather (plain and cluster multicast), float4 column reductions, warp-per-row-pair
# epilogue backward (plain and with the BatchNorm backward folded in)
graphs = [graph.synthetic_cell_graph(333, 8, 16, seed=20 + i, row_is_centre=True) for i in range(5)]
batch = graph.Batch.from_data_list(graphs).to(dev)
xp, sg = ops.build_sparse_graph(batch.x, batch.edge_index, batch.batch, batch.num_nodes_list, False)
ref = torch.empty_like(xp)
be.spmm(sg.rowptr, sg.col, sg.val, xp, None, sg.rs, ref, False)
for mc in (False, True):
    out = torch.empty_like(xp)
    be.sage_gather_ell(sg.ell, None, None, sg.ell_perm, sg.ell_width, sg.rowptr, sg.col, sg.val, xp, sg.num_graphs, sg.nmax,
                       out, tables_ready=True, two_per_sm=True, multicast=mc)
    torch.cuda.synchronize()
    check('gather two-per-SM multicast=%d' % mc, out, ref.cpu(), 1e-5)
rows, cin, cout = xp.shape[0], 20, 300
sa = torch.randn(rows, cin, device=dev)
w, bias = torch.randn(cin, cout, device=dev) / cin ** 0.5, torch.randn(cout, device=dev) * 0.1
y, gx = torch.empty(rows, cout, device=dev), torch.randn(rows, cout, device=dev)
rm, rv, mean, invstd = (torch.zeros(cout, device=dev) for _ in range(4))
be.sage_layer_fwd_wide(sa, w, bias, sg.mask, y, 0, rm, rv, 0.1, 1e-5, mean, invstd)
gamma, dg, db = torch.rand(cout, device=dev) + 0.5, torch.zeros(cout, device=dev), torch.zeros(cout, device=dev)
gy, dh1, dh2 = (torch.empty(rows, cout, device=dev) for _ in range(3))
be.bn_bwd(gx, y, mean, invstd, gamma, sg.mask, True, gy, dg, db)          # float4 sums + apply
be.sage_dh(gy, sa, w, bias, y, sg.mask, dh1, 0)
be.sage_dh_bn(gx, sa, w, bias, y, mean, invstd, gamma, dg, db, True, sg.mask, dh2, 0)
colsum = torch.zeros(cout, device=dev)
be.colsum(dh1, sg.mask, colsum)
torch.cuda.synchronize()
check('epilogue backward, BN folded', dh2, dh1.cpu(), 1e-4)
valid = (torch.arange(rows, device=dev) % sg.nmax) < sg.mask.num_nodes.repeat_interleave(sg.nmax)
check('float4 column sum', colsum, (dh1 * valid[:, None]).sum(0).cpu(), 1e-4)
print('sanitizer workload ok')
