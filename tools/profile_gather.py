"""Launch one kernel of the gather family a few times at the BASELINE shape, cold L2 (for ncu --set full).
VARIANT = ell (default) | layer | csr;  FEAT = 16."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cgc_net_b200 import _backend, graph, ops  # noqa: E402

b, n, k, feat = 64, 1800, 8, int(os.environ.get('FEAT', '16'))
be = _backend.get()
graphs = graph.synthetic_batch(8, n, k, feat, seed=1) * 8
batch = graph.Batch.from_data_list(graphs).to('cuda')
xp, g = ops.build_sparse_graph(batch.x, batch.edge_index, batch.batch, batch.num_nodes_list, False)
out = torch.empty_like(xp)
y = torch.empty(xp.shape[0], 20, device='cuda')
w, bias = torch.randn(feat, 20, device='cuda') * 0.2, torch.randn(20, device='cuda') * 0.1
rm, rv, mean, invstd = (torch.zeros(20, device='cuda') for _ in range(4))
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
variant = os.environ.get('VARIANT', 'ell')
for _ in range(3):
    flush.zero_()
    if variant == 'ell':
        be.sage_gather_ell(g.ell, None, None, g.ell_perm, g.ell_width, g.rowptr, g.col, None, xp, b, n, out,
                           tables_ready=True)
    elif variant == 'layer':
        be.sage_layer_ell_fwd(g.ell, None, g.ell_scale, g.ell_perm, g.ell_width, g.rowptr, g.col, None, xp, b, n, w,
                              bias, g.mask, 0, out, y, rm, rv, 0.1, 1e-5, mean, invstd, tables_ready=True)
    else:
        be.sage_gather_mean_fwd(g.rowptr, g.col, None, None, xp, b, n, out)
torch.cuda.synchronize()
print('done')
