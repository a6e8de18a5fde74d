"""Per-parameter gradient error of the CUDA path vs the float64 oracle (diagnostic, GPU box)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch  # noqa: E402

import helpers  # noqa: E402
from cgc_net_b200 import _backend, graph  # noqa: E402

graphs = graph.synthetic_batch(3, 1800, 8, 16, seed=2)
batch = graph.Batch.from_data_list(graphs)
ours, orc = helpers.build_pair((11404, 16, 20, 20, True, True, 20, 3, 0.1, [50]), load_data_sparse=True)
orc.train()
_, lossr = orc(batch)
lossr.backward()
_, _, truth = helpers.oracle_fp64(orc, batch)
ours = ours.cuda().train()
be = _backend.get()
for prec in (0, -1):
    be.gemm_precision = prec
    ours.zero_grad()
    lo, loss = ours(batch.to('cuda'))
    loss.backward()
    torch.cuda.synchronize()
    rows = []
    og = dict(orc.named_parameters())
    for k, p in ours.named_parameters():
        t = truth[k]
        sc = max(float(t.abs().max()), 1e-3)
        rows.append((float((p.grad.cpu().double() - t).abs().max()) / sc,
                     float((og[k].grad.double() - t).abs().max()) / sc, sc, k))
    rows.sort(reverse=True)
    print('--- gemm precision', prec, 'loss', float(loss))
    for r in rows[:8]:
        print('ours %.2e  oracle32 %.2e  scale %.2e  %s' % r)
