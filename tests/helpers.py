"""Shared builders for the parity tests."""
import torch

from cgc_net_b200 import graph, network
from oracle import restatement

GOLDEN_DIR = __import__('os').path.join(__import__('os').path.dirname(__file__), 'golden')


def make_graphs(sizes, degrees, feat=16, seed=0, messy=True):
    """Synthetic cell graphs; ``messy`` adds duplicate edges, drops a node's edges (isolated row)
    and alternates the edge direction convention, like the property cases of SURVEY section 4."""
    out = []
    for i, (n, k) in enumerate(zip(sizes, degrees)):
        d = graph.synthetic_cell_graph(n, k, feat, seed=seed * 1000 + i, row_is_centre=(i % 2 == 0))
        if messy and n > 4:
            ei = d.edge_index
            keep = ei[0] != 1                       # node 1 becomes an isolated row
            ei = ei[:, keep]
            ei = torch.cat([ei, ei[:, :5]], dim=1)  # duplicates
            perm = torch.randperm(ei.shape[1], generator=torch.Generator().manual_seed(i))
            d.edge_index = ei[:, perm]              # unsorted COO
        out.append(d)
    return out


def encoder_args(max_nodes=400, feat=16, hidden=20, ratio=0.1):
    return (max_nodes, feat, hidden, hidden, True, True, hidden, 3, ratio, [50])


def build_pair(args, seed=1, **kw):
    """(CUDA-path module, oracle) with identical parameters."""
    torch.manual_seed(seed)
    ours = network.SoftPoolingGcnEncoder(*args, **kw)
    orc = restatement.OracleEncoder(*args, **kw)
    orc.load_state_dict(ours.state_dict())
    return ours, orc


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def oracle_fp64(orc, batch):
    """Run a deep copy of the oracle in float64 (training mode) -> (logits, loss, {name: grad}).

    fp32 implementations that order their sums differently agree with each other only up to
    their own rounding (the fp32 reference itself is ~5e-5 off float64 on the most cancelling
    gradients), so gradient tolerances are asserted against this float64 evaluation."""
    import copy
    m = copy.deepcopy(orc).double()
    m.train()
    m.zero_grad()
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        if isinstance(batch, (tuple, list)):
            data = tuple(t.double() if torch.is_tensor(t) and t.is_floating_point() else t for t in batch)
        else:
            data = graph.Batch(batch.x.double(), batch.edge_index, batch.batch, batch.y, batch.num_nodes_list)
        logits, loss = m(data)
        loss.backward()
    finally:
        torch.set_default_dtype(prev)
    return logits.detach(), loss.detach(), {k: p.grad for k, p in m.named_parameters()}


def assert_grads_close(module, truth, ref32=None, tol=1e-4, floor=1e-3, slack=5.0):
    """Per parameter tensor: max|g - truth| <= tol * max(max|truth|, floor), where ``truth`` is the
    float64 oracle.  Gradients that are the residue of heavy cancellation (tiny pool stages) carry
    an absolute fp32 rounding error that no fp32 implementation can beat; for those the bound is
    ``slack`` x the fp32 oracle's OWN error against float64 (``ref32``), i.e. "as exact as the
    reference's arithmetic"."""
    for k, p in module.named_parameters():
        ref = truth[k].to(p.grad.device)
        err = float((p.grad.double() - ref).abs().max())
        bound = tol * max(float(ref.abs().max()), floor)
        if ref32 is not None:
            bound = max(bound, slack * float((ref32[k].double().to(ref.device) - ref).abs().max()))
        assert err <= bound, '%s: |dgrad| %.3e > %.3e' % (k, err, bound)
