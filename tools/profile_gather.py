"""Launch the SAGE gather kernel a few times at the BASELINE shape (for ncu --set full)."""
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
variant = os.environ.get('VARIANT', 'ell')
for _ in range(3):
    flush.zero_()
    if variant == 'ell':
        be.sage_gather_ell_fwd(g.ell, None, g.ell_scale, g.ell_perm, g.ell_width, g.rowptr, g.col, None, xp, b, n, out)
    else:
        be.sage_gather_mean_fwd(g.rowptr, g.col, None, None, xp, b, n, out)
torch.cuda.synchronize()
print('done')
