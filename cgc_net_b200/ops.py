"""Autograd boundary of the CUDA hot path.

Every numerical step here is ONE call into ``libcgc_sm100.so`` through ``_backend`` (the
C-ABI of ``include/cgc_sm100.h``); torch only allocates buffers and chains the
``torch.autograd.Function``s.  There is no CPU or eager fallback: ``_backend.get()``
raises if the library is missing.

Node features live in the PADDED node layout ``[B * Nmax, C]`` (row ``b * Nmax + i`` is
node ``i`` of graph ``b``): the reference normalises BatchNorm over all ``B * Nmax`` rows
and max-pools over all ``Nmax`` rows (``model/network.py:101-107,264``), so padded rows are
part of the arithmetic.  The stage-1 adjacency never becomes dense: it is a CSR (forward)
plus CSC (backward) pair over the padded rows with int32 indices.

Mapping to the reference:

* ``GnnModuleFunction``   = ``GNN_Module.forward`` (network.py:109-125) with PyG 1.2.1
  ``DenseSAGEConv.forward`` inlined (gather / mean / W / bias / L2-normalise / mask),
  ReLU, training-mode ``BatchNorm1d``, concat, optional ``lin``.
* ``DiffPoolFunction``    = ``_diff_pool`` (network.py:194-208).
* ``ColMaxFunction``      = ``torch.max(embed_feature, dim=1)`` (network.py:264,275,284).
* ``DenseRenormFunction`` = ``_re_norm_adj`` on a dense soft adjacency (network.py:183-191).
* ``build_sparse_graph``  = ``_sparse_to_dense_input`` (network.py:237-243) +
  ``construct_mask`` (:172-180) + stage-1 ``_re_norm_adj`` folded into edge weights.
"""
import torch
from torch.autograd import Function

from . import _backend

RENORM_P = 0.4       # model/network.py:260,271,280
BUILD_TRANSPOSED_ELL = False     # see build_sparse_graph
FOLD_BN_BACKWARD = True          # cgc_sage_dh_bn (BatchNorm backward formed inside the wide epilogue backward,
                                 # tests/test_gpu_ops.py::test_sage_dh_bn).  At the BASELINE shape (tools/epi_lab.py):
                                 # sums 186 us + folded epilogue 811 us = 997 us, against sums + apply 489 us +
                                 # epilogue 695 us = 1184 us with the warp-per-row-pair kernel (the register-capped
                                 # tile kernel of round 2a lost 479 us by folding and kept it off)
BUILD_BACKWARD_ELL = False       # ELL16 table of A^T for the narrow CSC gathers of the backward: measured 26.6 us per
                                 # gather (pitch 16, weighted, 7 index batches per thread) + 65 us to build the table vs
                                 # 24.9 us for the generic CSC gather -> off
ACTIVATIONS = {'relu': 0, 'elu': 1, 'leakyrelu': 2}     # GNN_Module._activation (network.py:82-90) -> CGC_ACT_*
BN_EPS = 1e-5
BN_MOMENTUM = 0.1

# Pass budget of the large tensor-core contractions: site -> ``precision`` argument of
# ``cgc_gemm`` (None = the backend's default, 3xTF32; 2 = one TF32 pass on round-to-nearest operands).  Only
# consulted when the backend is in its default mode: ``gemm_precision = 0`` (exact-fp32 test mode) overrides it.
# Measured at full size (tools/pass_budget.py, profiles/pass_budget_r2.txt): error of the single-pass run against
# the all-3xTF32 run, as a fraction of each gradient tensor's largest entry.
PASS_BUDGET = {
    'lin_fwd': None,        # lin of GNN_Module feeds softmax / arg-max(S): 52 cluster indices flip on one pass
    'lin_dw': 2,            # dW of lin: 5.6e-5
    'lin_dx': 2,            # dX of lin: 6.8e-5
    'pool_stx': None,       # X' = S^T X: logits move 4.5e-3 on one pass
    'pool_stas': None,      # A' = S^T (A S): fine at full size (logits 5e-7, gradients 6.4e-5, no index flips) but at
                            # small batches the perturbed A' moves stage-2 gradients past 1e-3 (smoke: 1.4e-3)
    'pool_gs': None,        # dS = P gA'^T + X gX'^T: 1.6e-3 on the small pool-2 tensors
    'pool_gp': 2,           # dP = S gA': 1.5e-5
    'pool_gx': None, 'pool_gadj': None,     # skinny / output-bound: nothing to gain
    'dense_agg': None,      # A' x on the dense stages: 4 index flips, 2.8e-2
    'dense_agg_t': None,    # no time to gain
    'dense_dadj': None,     # dA' = gA x^T: 1.3e-2
}


def _prec(be, site):
    if getattr(be, 'gemm_precision', -1) != -1:
        return None
    return PASS_BUDGET.get(site)


def _f32(*shape, device):
    return torch.empty(*shape, dtype=torch.float32, device=device)


def _mat(rows, cols, device):
    """``[rows, cols]`` fp32 matrix whose leading dimension is padded to a multiple of 4 floats
    (16 bytes): the alignment the TMA / tcgen05 contraction path needs (e.g. 114 clusters -> ld 116)."""
    ld = (cols + 3) // 4 * 4
    buf = torch.empty(rows, ld, dtype=torch.float32, device=device)
    return buf if ld == cols else buf[:, :cols]


class RowMask(object):
    """Validity of padded rows: row ``r`` is real iff ``r % nmax < num_nodes[r // nmax]``."""

    def __init__(self, num_nodes, nmax):
        self.num_nodes, self.nmax = num_nodes, int(nmax)


class SparseGraph(object):
    """Stage-1 adjacency over padded rows: CSR for ``A x``, CSC for ``A^T g``.

    ``val`` / ``val_t`` are the pooling-adjacency weights (``None`` = binary 0/1 adjacency;
    with ``norm_adj`` the ``_re_norm_adj`` values).  ``rs[i] = 1 / max(rowsum_i, 1)`` is the
    mean normaliser of DenseSAGEConv (``adj.sum(-1).clamp(min=1)``).
    """

    def __init__(self):
        self.num_graphs = self.nmax = self.rows = 0
        self.num_nodes = None
        self.rowptr = self.col = self.val = None
        self.colptr = self.row = self.val_t = None
        self.rs = None
        self.mask = None
        self.ell = self.ell_val = self.ell_scale = self.ell_perm = None   # row-padded CSR (cgc_csr_to_ell16)
        self.ell_t = self.ell_t_val = self.ell_t_scale = self.ell_t_perm = None   # the same for A^T (from the CSC)
        self.ell_t_width = 0
        self.ell_width = 0
        # table of A^T with the mean normaliser of the SOURCE row folded into the edge weights (rs[i] * A[i, j] at
        # table row j): the CSC gathers of the narrow-layer backward, dX = A^T (rs * dsa)
        self.ell_b = self.ell_b_val = self.ell_b_scale = self.ell_b_perm = self.val_b = None
        self.ell_b_width = 0
        # True until the first kernel that reads the ELL16 tables has been enqueued: that consumer directly follows
        # the table builder in the stream, so it must not read the tables in its programmatic-launch prologue
        self.tables_fresh = True

    def tables_ready(self):
        """For the ``tables_ready`` flag of the ELL16 kernels: False exactly once after the build."""
        fresh, self.tables_fresh = self.tables_fresh, False
        return not fresh


def build_sparse_graph(x, edge_index, batch, num_nodes_list=None, norm_adj=False, max_degree=None, self_loops=None):
    """COO ``edge_index`` + sorted ``batch`` -> (padded ``x``, ``SparseGraph``).

    Replaces ``to_dense_adj`` (model/utils.py:3-37: ``adj[b, src, dst] = 1``, duplicates
    collapse) and ``to_dense_batch`` without materialising ``[B, Nmax, Nmax]``.
    ``max_degree`` / ``self_loops``: optional HOST facts from the loader (``graph.Batch``) -- the largest number of
    COO entries of one source row, and whether every node has a self loop -- used to pick the index-table pitch
    without a device sync (pitch 8 only when no row can exceed it; otherwise 16, rows above that walk the CSR).
    """
    be = _backend.get()
    dev = x.device
    n_total = x.shape[0]
    if num_nodes_list is not None:
        num_graphs, nmax = len(num_nodes_list), int(max(num_nodes_list))
        num_nodes = torch.tensor(num_nodes_list, dtype=torch.int32).to(dev, non_blocking=True)
    else:                                   # same host syncs as model/utils.py:17,21
        num_graphs = int(batch[-1].item()) + 1
        num_nodes = torch.zeros(num_graphs, dtype=torch.int32, device=dev)
        be.count_nodes(batch, num_nodes)
        nmax = int(num_nodes.max().item())
    start = torch.zeros(num_graphs + 1, dtype=torch.int32, device=dev)
    start[1:] = torch.cumsum(num_nodes, 0)
    g = SparseGraph()
    g.num_graphs, g.nmax, g.rows = num_graphs, nmax, num_graphs * nmax
    g.num_nodes = num_nodes
    g.mask = RowMask(num_nodes, nmax)
    rows = g.rows
    xp = _f32(rows, x.shape[1], device=dev)
    be.pad_nodes(x.contiguous(), batch, start, nmax, xp)

    n_edges = edge_index.shape[1]
    cap = n_edges + (n_total if norm_adj else 0)     # renorm forces the diagonal (network.py:188)
    g.rowptr = torch.empty(rows + 1, dtype=torch.int32, device=dev)
    g.col = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
    work = torch.empty(be.coo_to_csr_workspace(cap, rows), dtype=torch.int32, device=dev)
    be.coo_to_csr(edge_index.contiguous(), batch, start, num_nodes, nmax, num_graphs, n_total,
                  bool(norm_adj), g.rowptr, g.col, work)
    g.rs = _f32(rows, device=dev)
    if norm_adj:
        g.val = _f32(max(cap, 1), device=dev)
        be.edge_renorm(g.rowptr, g.col, RENORM_P, g.val, g.rs)
    else:
        be.row_inv_degree(g.rowptr, g.rs)
    g.colptr = torch.empty(rows + 1, dtype=torch.int32, device=dev)
    g.row = torch.zeros(max(cap, 1), dtype=torch.int32, device=dev)
    g.val_t = _f32(max(cap, 1), device=dev) if norm_adj else None
    be.csr_transpose(g.rowptr, g.col, g.val, g.colptr, g.row, g.val_t, work)
    # Fixed-pitch index table of the shared-memory gathers (layer-1 SAGE gather when the feature width allows it,
    # wide A*S of the DiffPool forward always): pitch 8 when the mean row degree allows it, else 16; rows above the
    # pitch are flagged and walk the CSR.  The slot-parity / row-pairing order serves the 64-byte shared-memory rows
    # both kernels use.
    real_rows = max(n_total, 1)
    if nmax + 1 <= 0xFFF and cap <= 16 * real_rows and (nmax + 2) * 64 + 1024 <= 227 * 1024:
        if max_degree is not None:
            # renorm forces the diagonal: one more entry per row unless every node already has its self loop
            worst = int(max_degree) + (1 if norm_adj and not self_loops else 0)
            g.ell_width = 8 if worst <= 8 else 16
        else:
            g.ell_width = 8 if cap <= 8 * real_rows else 16
        trows = be.ell16_rows(num_graphs, nmax)
        g.ell = torch.empty(trows * g.ell_width, dtype=torch.int16, device=dev)
        g.ell_val = _f32(trows * g.ell_width, device=dev) if norm_adj else None
        g.ell_scale = _f32(trows, device=dev)
        g.ell_perm = torch.empty(trows // 8, dtype=torch.int32, device=dev)
        be.csr_to_ell16(g.rowptr, g.col, g.val, g.rs, num_graphs, nmax, g.ell_width, True, g.ell, g.ell_val,
                        g.ell_scale, g.ell_perm)
        if BUILD_BACKWARD_ELL and torch.is_grad_enabled():
            # weights of the transposed edges: rs of the edge's source row (times the renormalised value); the unused
            # tail of g.row is zero-filled (valid indices)
            g.val_b = torch.index_select(g.rs, 0, g.row)
            if norm_adj:
                g.val_b = g.val_b * g.val_t
            g.ell_b_width = 16                      # in-degrees spread around the mean: pitch 16, the rest walks the CSC
            g.ell_b = torch.empty(trows * 16, dtype=torch.int16, device=dev)
            g.ell_b_val = _f32(trows * 16, device=dev)
            g.ell_b_scale = _f32(trows, device=dev)
            g.ell_b_perm = torch.empty(trows // 8, dtype=torch.int32, device=dev)
            be.csr_to_ell16(g.colptr, g.row, g.val_b, None, num_graphs, nmax, 16, True, g.ell_b, g.ell_b_val,
                            g.ell_b_scale, g.ell_b_perm)
        if BUILD_TRANSPOSED_ELL:
            # table of the transposed adjacency (wide gather of the DiffPool backward, A^T gP).  In-degrees vary more
            # than out-degrees, the pitch has to be 16 and half of the slots are padding: measured 620 us vs 425 us
            # for the generic CSC gather at the BASELINE shape, so it is off by default.
            g.ell_t_width = 8 if 2 * cap <= 8 * real_rows else 16
            g.ell_t = torch.empty(trows * g.ell_t_width, dtype=torch.int16, device=dev)
            g.ell_t_val = _f32(trows * g.ell_t_width, device=dev) if norm_adj else None
            g.ell_t_scale = _f32(trows, device=dev)
            g.ell_t_perm = torch.empty(trows // 8, dtype=torch.int32, device=dev)
            be.csr_to_ell16(g.colptr, g.row, g.val_t, None, num_graphs, nmax, g.ell_t_width, True, g.ell_t,
                            g.ell_t_val, g.ell_t_scale, g.ell_t_perm)
    return xp, g


# ----------------------------------------------------------------------------- GNN module
class GnnConfig(object):
    """Static description of one ``GNN_Module`` call (not a tensor; passed through autograd)."""

    def __init__(self, graph=None, num_graphs=0, nmax=0, mask=None, training=True, has_bn=True,
                 has_lin=False, bn_buffers=None, need_adj_grad=False, act=0, agg_cache=None):
        self.graph = graph              # SparseGraph, or None when the adjacency is dense
        self.num_graphs, self.nmax = num_graphs, nmax
        self.mask = mask                # RowMask or None
        self.training, self.has_bn, self.has_lin = training, has_bn, has_lin
        self.bn_buffers = bn_buffers    # [(running_mean, running_var)] * 3 (updated in place)
        self.need_adj_grad = need_adj_grad
        # dense stages: {(adjacency, layer-1 input) -> aggregated input}.  GCN_embed_k and GCN_pool_k aggregate the
        # SAME x over the SAME adjacency in their first layer (network.py:270-276): the second module reuses the
        # first one's A'x instead of reading the 333 MB adjacency again
        self.agg_cache = agg_cache
        self.act = act                  # ACTIVATIONS[name]: 0 relu, 1 elu, 2 leakyrelu (network.py:82-90)


def _agg(be, cfg, adj, rs, x_prev, out):
    """Mean-aggregated neighbours ``sa = rs * (A x)`` (DenseSAGEConv: matmul(adj, x) / deg)."""
    if cfg.graph is not None:
        g = cfg.graph
        feat = x_prev.shape[1]
        dense_io = x_prev.is_contiguous() and out.is_contiguous() and feat % 4 == 0 and feat <= 32
        if _ell_ok(g, x_prev, out) and rs is g.rs:
            # fixed-pitch index table (rs baked in) + TMA-staged feature tile (the roofline kernel); column slices
            # of the concat buffer (layers 2-3) are staged by 2-D TMA boxes
            # binary adjacency: no scale vector (the kernel derives the mean 1 / max(deg, 1) from the row's slots)
            be.sage_gather_ell(g.ell, g.ell_val, g.ell_scale if g.ell_val is not None else None, g.ell_perm,
                               g.ell_width, g.rowptr, g.col, g.val, x_prev, g.num_graphs, g.nmax, out,
                               tables_ready=g.tables_ready())
        elif dense_io and g.nmax * feat * 4 + 12288 <= 220 * 1024:
            # whole-graph feature tile fits in shared memory: TMA-staged gather (the roofline kernel)
            be.sage_gather_mean_fwd(g.rowptr, g.col, g.val, rs, x_prev, g.num_graphs, g.nmax, out)
        else:
            be.spmm(g.rowptr, g.col, g.val, x_prev, None, rs, out, False)
    else:
        b, n = cfg.num_graphs, cfg.nmax
        be.gemm(adj, _as3d(x_prev, b, n), out.view(b, n, -1), False, False, row_scale=rs,
                precision=_prec(be, 'dense_agg'))


def _ell_ok(g, x, out):
    """The shared-memory ELL16 gather serves feature widths that are multiples of 4 up to 32, 16-byte aligned rows
    (leading dimensions multiples of 4 floats) and tiles that fit: (round_up(nmax, box) rows) * feat * 4 <= 225 KB."""
    feat = x.shape[1]
    return (g.ell is not None and feat % 4 == 0 and feat <= 32 and x.stride(1) == 1 and out.stride(1) == 1
            and x.stride(0) % 4 == 0 and out.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0
            and out.data_ptr() % 16 == 0 and (g.nmax + 256) * feat * 4 <= 225 * 1024)


def _agg_t(be, cfg, adj, g_a, src_scale, out, accumulate):
    """``A^T (src_scale * g)`` into ``out`` (``+=`` when ``accumulate``)."""
    if cfg.graph is not None:
        g = cfg.graph
        if g.ell_b is not None and src_scale is g.rs and _ell_ok(g, g_a, out):
            # rs is folded into the table's edge weights; a column slice of the gradient buffer is a strided out
            be.sage_gather_ell(g.ell_b, g.ell_b_val, g.ell_b_scale, g.ell_b_perm, g.ell_b_width, g.colptr, g.row,
                               g.val_b, g_a, g.num_graphs, g.nmax, out, accumulate=accumulate, tables_ready=True)
        else:
            be.spmm(g.colptr, g.row, g.val_t, g_a, src_scale, None, out, accumulate)
    else:
        b, n = cfg.num_graphs, cfg.nmax
        be.gemm(adj, _as3d(g_a, b, n), _as3d(out, b, n), True, False,
                beta=1.0 if accumulate else 0.0, precision=_prec(be, 'dense_agg_t'))


def _aligned_weight(w):
    """``w`` itself when its rows are 16-byte aligned, else a copy with the leading dimension padded to 4 floats
    (lin of the second pooling module is [114, 154]: as stored it would keep the contraction off the TMA path)."""
    if w.stride(0) % 4 == 0 and w.data_ptr() % 16 == 0:
        return w
    wp = _mat(w.shape[0], w.shape[1], w.device)
    wp.copy_(w)
    return wp


def _as3d(t, b, n):
    """View a (possibly column-sliced) ``[b*n, c]`` matrix as ``[b, n, c]`` without copying."""
    return t.as_strided((b, n, t.shape[1]), (n * t.stride(0), t.stride(0), 1), t.storage_offset())


class GnnModuleFunction(Function):
    """``GNN_Module.forward`` (network.py:109-125), forward and hand-written backward.

    Inputs: ``x0 [M, Cin]``; ``adj`` = dense ``[B, N, N]`` adjacency or ``None`` (sparse stage 1);
    ``rs [M]`` mean normaliser and (dense only) ``rsum [M]`` the raw row sums; then per conv
    layer ``weight [Cin, Cout]``, ``bias``, ``bn.weight``, ``bn.bias``; finally
    ``lin.weight [E, Ctot]``, ``lin.bias`` (or ``None``s).
    Output: ``[M, Ctot]`` concat (``lin=False``) or ``[M, E]``.
    """

    @staticmethod
    def forward(ctx, cfg, x0, adj, rs, rsum, *params):
        be = _backend.get()
        dev = x0.device
        m = x0.shape[0]
        conv = [params[4 * k: 4 * k + 4] for k in range(3)]
        w_lin, b_lin = params[12], params[13]
        couts = [c[0].shape[1] for c in conv]
        ctot = sum(couts)
        cat = _mat(m, ctot, dev)                     # leading dimension padded to 4 floats (TMA operand of lin)
        saved = []
        x_prev, off = x0, 0
        for k in range(3):
            w, b, gamma, beta = conv[k]
            a = _f32(m, w.shape[0], device=dev)          # sa = rs * (A x)
            y = _f32(m, couts[k], device=dev)
            dst = cat[:, off: off + couts[k]]
            mean = invstd = None
            if cfg.has_bn:
                mean = _f32(couts[k], device=dev)
                invstd = _f32(couts[k], device=dev)
                rmean, rvar = cfg.bn_buffers[k]
            g = cfg.graph
            if (g is not None and rs is g.rs and _ell_ok(g, x_prev, a)
                    and be.sage_layer_ell_supported(w.shape[0], couts[k], g.nmax)):
                # the whole layer up to the BatchNorm statistics in ONE kernel (gather + W + bias + L2 norm + mask
                # + activation + column partial sums)
                stats = cfg.has_bn and cfg.training
                be.sage_layer_ell_fwd(g.ell, g.ell_val, g.ell_scale if g.ell_val is not None else None, g.ell_perm,
                                      g.ell_width, g.rowptr, g.col, g.val, x_prev, g.num_graphs, g.nmax, w, b, cfg.mask, cfg.act, a, y,
                                      rmean if stats else None, rvar if stats else None, BN_MOMENTUM, BN_EPS,
                                      mean if stats else None, invstd if stats else None,
                                      tables_ready=g.tables_ready())
                if cfg.has_bn and not cfg.training:
                    be.bn_eval_stats(rmean, rvar, BN_EPS, mean, invstd)
                be.bn_apply(y, mean, invstd, gamma if cfg.has_bn else None,
                            beta if cfg.has_bn else None, cfg.mask, dst)
                saved += [a, y, mean, invstd]
                x_prev, off = dst, off + couts[k]
                continue
            key = (adj.data_ptr(), x_prev.data_ptr(), tuple(x_prev.shape)) if (k == 0 and g is None) else None
            if key is not None and cfg.agg_cache is not None and key in cfg.agg_cache:
                a = cfg.agg_cache[key]
            else:
                _agg(be, cfg, adj, rs, x_prev, a)
                if key is not None and cfg.agg_cache is not None:
                    cfg.agg_cache[key] = a
            if cfg.has_bn and cfg.training and couts[k] <= 32 and w.shape[0] <= 64:
                # narrow layer: epilogue and BatchNorm batch statistics in ONE pass over the rows
                be.sage_layer_fwd_narrow(a, w, b, cfg.mask, y, cfg.act, rmean, rvar, BN_MOMENTUM, BN_EPS, mean, invstd)
            elif cfg.has_bn and cfg.training and 32 < couts[k] <= 1280 and w.shape[0] <= 64:
                # wide layer (20 -> 1140): the statistics ride on the epilogue, no separate pass over y
                be.sage_layer_fwd_wide(a, w, b, cfg.mask, y, cfg.act, rmean, rvar, BN_MOMENTUM, BN_EPS, mean, invstd)
            else:
                be.sage_epilogue_fwd(a, w, b, cfg.mask, y, cfg.act)
                if cfg.has_bn and cfg.training:
                    be.bn_stats(y, rmean, rvar, BN_MOMENTUM, BN_EPS, mean, invstd)
                elif cfg.has_bn:
                    be.bn_eval_stats(rmean, rvar, BN_EPS, mean, invstd)
            be.bn_apply(y, mean, invstd, gamma if cfg.has_bn else None,
                        beta if cfg.has_bn else None, cfg.mask, dst)
            saved += [a, y, mean, invstd]
            x_prev, off = dst, off + couts[k]
        if cfg.has_lin:
            out = _f32(m, w_lin.shape[0], device=dev)
            be.gemm(cat, _aligned_weight(w_lin), out, False, True, bias=b_lin, rowmask=cfg.mask,
                    precision=_prec(be, 'lin_fwd'))
        else:
            out = cat
        ctx.cfg, ctx.couts = cfg, couts
        ctx.save_for_backward(x0, adj, rs, rsum, cat, *saved, *params)
        return out

    @staticmethod
    def backward(ctx, g_out):
        be = _backend.get()
        cfg, couts = ctx.cfg, ctx.couts
        t = ctx.saved_tensors
        x0, adj, rs, rsum, cat = t[0:5]
        saved, params = t[5:17], t[17:]
        conv = [params[4 * k: 4 * k + 4] for k in range(3)]
        w_lin, b_lin = params[12], params[13]
        dev = g_out.device
        m = cat.shape[0]
        if g_out.stride(1) != 1 or g_out.stride(0) % 4 != 0 or g_out.data_ptr() % 16 != 0:
            g_out = g_out.contiguous()
        grads = [None] * 14
        if cfg.has_lin:
            d_wl = torch.empty_like(w_lin)
            be.gemm(g_out, cat, d_wl, True, False, precision=_prec(be, 'lin_dw'))     # dY^T X (K = M; padded rows of X are 0)
            grads[12] = d_wl
            if b_lin is not None:
                d_bl = torch.empty_like(b_lin)
                be.colsum(g_out, cfg.mask, d_bl)
                grads[13] = d_bl
            g_cat = _mat(m, cat.shape[1], dev)
            be.gemm(g_out, _aligned_weight(w_lin), g_cat, False, False, rowmask=cfg.mask,
                    precision=_prec(be, 'lin_dx'))                                                        # (dY * mask) W
        else:
            g_cat = g_out.clone()                              # accumulated into below
        dense = cfg.graph is None
        want_adj = dense and ctx.needs_input_grad[2]
        d_adj = torch.empty_like(adj) if want_adj else None
        d_rs = torch.zeros_like(rs) if dense else None
        g_x0 = None
        offs = [0, couts[0], couts[0] + couts[1]]
        if want_adj:
            # dA = sum_k gA_k x_{k-1}^T as ONE contraction over the concatenated channels
            # [layer-1 | layer-2 | layer-3]: gA_k land in slices of ga_all, the layer inputs are x0 and
            # the first two blocks of the concat buffer.
            cin0 = x0.shape[1]
            goffs = [0, cin0, cin0 + couts[0]]
            ktot = cin0 + couts[0] + couts[1]
            ga_all = _f32(m, ktot, device=dev)
            xc_all = _f32(m, ktot, device=dev)
            be.bn_apply(x0, None, None, None, None, None, xc_all[:, :cin0])                       # strided copies
            be.bn_apply(cat[:, :couts[0] + couts[1]], None, None, None, None, None, xc_all[:, cin0:])
        for k in (2, 1, 0):
            w, b, gamma, beta = conv[k]
            a, y, mean, invstd = saved[4 * k: 4 * k + 4]
            gx = g_cat[:, offs[k]: offs[k] + couts[k]]
            gy = _f32(m, couts[k], device=dev)
            narrow = w.shape[1] <= 32 and w.shape[0] <= 64
            fold_bn = (FOLD_BN_BACKWARD and cfg.has_bn and not narrow and 32 < w.shape[1] <= 1280
                       and w.shape[0] <= 64)
            if cfg.has_bn:
                d_gamma, d_beta = torch.empty_like(gamma), torch.empty_like(beta)
                # wide layer: only the column sums here, the apply step is folded into the epilogue backward below
                be.bn_bwd(gx, y, mean, invstd, gamma, cfg.mask, cfg.training, None if fold_bn else gy, d_gamma, d_beta)
                grads[4 * k + 2], grads[4 * k + 3] = d_gamma, d_beta
            else:
                be.bn_apply(gx, None, None, None, None, cfg.mask, gy)     # gy = gx * mask
            d_w = torch.empty_like(w)
            d_b = torch.empty_like(b) if b is not None else None
            dsa = _f32(m, w.shape[0], device=dev)
            if narrow:
                # narrow layer: dh, dW, db and dsa = dh W^T in ONE pass over the rows
                be.sage_layer_bwd_narrow(gy, a, w, b, y, cfg.mask, dsa, d_w, d_b, cfg.act)
            else:
                if fold_bn:
                    be.sage_dh_bn(gx, a, w, b, y, mean, invstd, gamma, d_gamma, d_beta, cfg.training, cfg.mask, gy,
                                  cfg.act)                      # gy <- dh, BatchNorm backward formed on the fly
                else:
                    be.sage_dh(gy, a, w, b, y, cfg.mask, gy, cfg.act)   # gy <- dh = dL/d(sa W + b), in place
                be.gemm(a, gy, d_w, True, False)                # dW = sa^T dh   (K = M)
                if d_b is not None:
                    be.colsum(gy, None, d_b)
                be.gemm(gy, w, dsa, False, True)                # dsa = dh W^T
            grads[4 * k], grads[4 * k + 1] = d_w, d_b
            x_prev = x0 if k == 0 else cat[:, offs[k - 1]: offs[k - 1] + couts[k - 1]]
            if dense:
                if want_adj:
                    ga = ga_all[:, goffs[k]: goffs[k] + w.shape[0]]
                else:
                    ga = _f32(m, w.shape[0], device=dev)
                be.scale_rows_bwd(dsa, a, rs, ga, d_rs)         # ga = rs*dsa ; d_rs += dsa.sa/rs
                src_scale = None
            else:
                ga, src_scale = dsa, rs                         # rs applied inside the CSC gather
            if k > 0:
                _agg_t(be, cfg, adj, ga, src_scale,
                       g_cat[:, offs[k - 1]: offs[k - 1] + couts[k - 1]], True)
            elif ctx.needs_input_grad[1]:
                g_x0 = _f32(m, x0.shape[1], device=dev)
                _agg_t(be, cfg, adj, ga, src_scale, g_x0, False)
        if want_adj:
            b_, n_ = cfg.num_graphs, cfg.nmax
            be.gemm(ga_all.view(b_, n_, -1), xc_all.view(b_, n_, -1), d_adj, False, True,
                    precision=_prec(be, 'dense_dadj'))                                       # dA = gA x^T
            # rs_i = 1 / clamp(rsum_i, 1)  ->  dA_ij += -[rsum_i >= 1] d_rs_i / rsum_i^2   (all j)
            be.add_rs_grad(d_adj, rsum, d_rs)
        return (None, g_x0, d_adj, None, None) + tuple(grads)


# ----------------------------------------------------------------------------- GIN variant (composable pieces)
class NeighbourSumFunction(Function):
    """``torch.matmul(adj, x)`` of DenseGINConv (PyG 1.2.1; reference call site network.py:99,114): the plain
    neighbour SUM (no mean, no self term since the reference passes ``add_loop=False``).  Sparse stage 1: CSR
    gather forward, CSC gather backward; dense stages: batched contraction with a gradient for the adjacency."""

    @staticmethod
    def forward(ctx, cfg, x, adj):
        be = _backend.get()
        out = _f32(x.shape[0], x.shape[1], device=x.device)
        if cfg.graph is not None:
            g = cfg.graph
            be.spmm(g.rowptr, g.col, g.val, x, None, None, out, False)
        else:
            b, n = cfg.num_graphs, cfg.nmax
            be.gemm(adj, _as3d(x, b, n), out.view(b, n, -1), False, False)
        ctx.cfg = cfg
        ctx.save_for_backward(x, adj)
        return out

    @staticmethod
    def backward(ctx, g_out):
        be = _backend.get()
        cfg = ctx.cfg
        x, adj = ctx.saved_tensors
        g_out = g_out.contiguous()
        gx = g_adj = None
        if cfg.graph is not None:
            if ctx.needs_input_grad[1]:
                g = cfg.graph
                gx = _f32(x.shape[0], x.shape[1], device=x.device)
                be.spmm(g.colptr, g.row, g.val_t, g_out, None, None, gx, False)
        else:
            b, n = cfg.num_graphs, cfg.nmax
            if ctx.needs_input_grad[1]:
                gx = _f32(x.shape[0], x.shape[1], device=x.device)
                be.gemm(adj, _as3d(g_out, b, n), gx.view(b, n, -1), True, False)           # A^T g
            if ctx.needs_input_grad[2]:
                g_adj = torch.empty_like(adj)
                be.gemm(_as3d(g_out, b, n), _as3d(x.contiguous(), b, n), g_adj, False, True)   # g x^T
        return None, gx, g_adj


class LinearFunction(Function):
    """``nn.Linear`` of the GIN MLP (network.py:97-98): ``y = (x W^T + b) [* mask]``."""

    @staticmethod
    def forward(ctx, x, weight, bias, mask):
        be = _backend.get()
        x = x.contiguous()
        y = _f32(x.shape[0], weight.shape[0], device=x.device)
        be.gemm(x, weight, y, False, True, bias=bias, rowmask=mask)
        ctx.mask = mask
        ctx.save_for_backward(x, weight, bias)
        return y

    @staticmethod
    def backward(ctx, gy):
        be = _backend.get()
        x, weight, bias = ctx.saved_tensors
        gy = gy.contiguous()
        if ctx.mask is not None:
            gm = torch.empty_like(gy)
            be.bn_apply(gy, None, None, None, None, ctx.mask, gm)      # gy * mask
            gy = gm
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = _f32(x.shape[0], x.shape[1], device=x.device)
            be.gemm(gy, weight, gx, False, False)
        if ctx.needs_input_grad[1]:
            gw = torch.empty_like(weight)
            be.gemm(gy, x, gw, True, False)
        if bias is not None and ctx.needs_input_grad[2]:
            gb = torch.empty_like(bias)
            be.colsum(gy, None, gb)
        return gx, gw, gb, None


class ActFunction(Function):
    """relu / elu / leakyrelu (``GNN_Module._activation``, network.py:82-90); gradient from the output."""

    @staticmethod
    def forward(ctx, x, act):
        be = _backend.get()
        y = torch.empty_like(x, memory_format=torch.contiguous_format)
        be.act_fwd(x.contiguous(), act, y)
        ctx.act = act
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, gy):
        be = _backend.get()
        y, = ctx.saved_tensors
        gx = torch.empty_like(y)
        be.act_bwd(gy.contiguous(), y, ctx.act, gx)
        return gx, None


class BatchNormFunction(Function):
    """``nn.BatchNorm1d`` over all padded rows (network.py:101-107), output multiplied by the row mask
    (network.py:119-120 applies it to the concat; padded rows never reach real rows before that)."""

    @staticmethod
    def forward(ctx, y, gamma, beta, rmean, rvar, training, mask):
        be = _backend.get()
        y = y.contiguous()
        c = y.shape[1]
        mean, invstd = _f32(c, device=y.device), _f32(c, device=y.device)
        if training:
            be.bn_stats(y, rmean, rvar, BN_MOMENTUM, BN_EPS, mean, invstd)
        else:
            be.bn_eval_stats(rmean, rvar, BN_EPS, mean, invstd)
        out = torch.empty_like(y)
        be.bn_apply(y, mean, invstd, gamma, beta, mask, out)
        ctx.training, ctx.mask = training, mask
        ctx.save_for_backward(y, mean, invstd, gamma)
        return out

    @staticmethod
    def backward(ctx, g):
        be = _backend.get()
        y, mean, invstd, gamma = ctx.saved_tensors
        gy = torch.empty_like(y)
        d_gamma, d_beta = torch.empty_like(gamma), torch.empty_like(gamma)
        be.bn_bwd(g.contiguous(), y, mean, invstd, gamma, ctx.mask, ctx.training, gy, d_gamma, d_beta)
        return gy, d_gamma, d_beta, None, None, None, None


# ----------------------------------------------------------------------------- DenseJK
class DenseJKFunction(Function):
    """``DenseJK.forward`` (network.py:36-52) as one fused kernel per direction of autograd: forward writes a
    tape in training mode, backward turns it into ``d feats`` plus gate / score gradients whose contractions over
    the nodes (the weight and bias gradients) run on the library GEMM.

    ``lstm_params`` = (weight_ih_l0, weight_hh_l0, bias_ih_l0, bias_hh_l0, weight_ih_l0_reverse, ...)."""

    @staticmethod
    def forward(ctx, feats, channels, layers, att_w, att_b, *lstm_params):
        be = _backend.get()
        feats = feats.contiguous()
        rows = feats.shape[0]
        out = _f32(rows, channels, device=feats.device)
        need_tape = any(ctx.needs_input_grad)           # inference: no tape is written
        tape = _f32(be.jk_tape_floats(rows, channels, layers), device=feats.device) if need_tape else None
        be.jk_lstm_fwd(feats, channels, layers, lstm_params, att_w, att_b, out, tape)
        ctx.dims = (channels, layers)
        ctx.tape = tape
        ctx.save_for_backward(feats, att_w, *lstm_params)
        return out

    @staticmethod
    def backward(ctx, g_out):
        be = _backend.get()
        channels, layers = ctx.dims
        feats, att_w = ctx.saved_tensors[:2]
        lstm_params = ctx.saved_tensors[2:]
        tape = ctx.tape
        dev = feats.device
        rows = feats.shape[0]
        hid = channels * layers // 2
        g_feats = _f32(rows, layers * channels, device=dev)
        dgates = _f32(2, layers, 4 * hid, rows, device=dev)
        dscore = _f32(layers, rows, device=dev)
        be.jk_lstm_bwd(g_out.contiguous(), feats, channels, layers, lstm_params, att_w, tape, g_feats, dgates, dscore)
        n_g = 2 * layers * 5 * hid * rows
        hs = tape[n_g:n_g + 2 * layers * hid * rows].view(2, layers, hid, rows)
        ones = torch.ones(rows, 1, dtype=torch.float32, device=dev)
        grads = []
        d_att_w = torch.empty(1, 2 * hid, dtype=torch.float32, device=dev)
        for d in range(2):
            order = list(range(layers)) if d == 0 else list(range(layers - 1, -1, -1))
            d_wih = _f32(4 * hid, channels, device=dev)
            d_whh = _f32(4 * hid, hid, device=dev)
            d_b = _f32(4 * hid, 1, device=dev)
            for i, t in enumerate(order):
                first = i == 0
                x_t = feats[:, t * channels:(t + 1) * channels]
                be.gemm(dgates[d, t], x_t, d_wih, False, False, beta=0.0 if first else 1.0)     # K = rows
                be.gemm(dgates[d, t], ones, d_b, False, False, beta=0.0 if first else 1.0)
                if i > 0:                                    # h_prev of the first computed step is zero
                    be.gemm(dgates[d, t], hs[d, order[i - 1]], d_whh, False, True, beta=0.0 if i == 1 else 1.0)
                be.gemm(dscore[t:t + 1], hs[d, t], d_att_w[:, d * hid:(d + 1) * hid], False, True,
                        beta=0.0 if first else 1.0)
            if layers == 1:
                d_whh.zero_()
            grads += [d_wih, d_whh, d_b.view(-1), d_b.view(-1).clone()]
        d_att_b = torch.empty(1, dtype=torch.float32, device=dev)
        be.colsum(dscore.view(-1, 1), None, d_att_b)
        return (g_feats, None, None, d_att_w, d_att_b) + tuple(grads)


def _wide_ell_ok(g, table, x, out):
    """The shared-memory wide gather needs an ELL16 table, 16-byte aligned rows and a slice tile that fits."""
    return (table is not None and x.shape[1] >= 64 and x.shape[1] % 4 == 0 and x.stride(0) % 4 == 0
            and out.stride(0) % 4 == 0 and (g.nmax + 2) * 64 <= 226 * 1024)


# ----------------------------------------------------------------------------- diff-pool
class DiffPoolFunction(Function):
    """``_diff_pool`` (network.py:194-208): ``S = softmax(Z) * mask``; ``X' = S^T X``; ``A' = S^T (A S)``.

    Returns ``(X' [B,C,F], A' [B,C,C], S_full or None, argmax or None)``; the last two are the
    un-masked softmax (``collect_assign``, network.py:201-202) and the cluster index per node.
    """

    @staticmethod
    def forward(ctx, cfg, z, x, adj, collect):
        be = _backend.get()
        dev = z.device
        m, c = z.shape
        b, n = cfg.num_graphs, cfg.nmax
        s = _mat(m, c, dev)                      # leading dimension padded to 4 floats (TMA operand)
        s_full = _f32(m, c, device=dev) if collect else None
        arg = torch.empty(m, dtype=torch.int32, device=dev) if collect else None
        be.softmax_fwd(z.contiguous(), cfg.mask, s, s_full, arg)
        x = x.contiguous()
        xp = _f32(b, c, x.shape[1], device=dev)
        be.gemm(_as3d(s, b, n), x.view(b, n, -1), xp, True, False, precision=_prec(be, 'pool_stx'))
        p = _mat(m, c, dev)
        if cfg.graph is not None:
            g = cfg.graph
            if _wide_ell_ok(g, g.ell, s, p):
                be.gather_wide_ell(g.ell, g.ell_val, None, g.ell_perm, g.ell_width, g.rowptr, g.col, g.val, s,
                                   g.num_graphs, g.nmax, p, False)
            else:
                be.spmm(g.rowptr, g.col, g.val, s, None, None, p, False)
        else:
            be.gemm(adj, _as3d(s, b, n), _as3d(p, b, n), False, False)
        ap = _f32(b, c, c, device=dev)
        be.gemm(_as3d(s, b, n), _as3d(p, b, n), ap, True, False, precision=_prec(be, 'pool_stas'))
        ctx.cfg = cfg
        ctx.has_adj = adj is not None
        # s / p are intermediates (never returned), x / adj inputs: saved through autograd so that in-place edits
        # are caught and nothing forms a tensor -> grad_fn -> ctx -> tensor cycle
        ctx.save_for_backward(s, p, x, *([adj] if adj is not None else []))
        ctx.want_adj = adj is not None and ctx.needs_input_grad[3]
        outs = (xp, ap, s_full, arg)
        ctx.mark_non_differentiable(*[t for t in (s_full, arg) if t is not None])
        return outs

    @staticmethod
    def backward(ctx, g_xp, g_ap, _gs_full, _garg):
        be = _backend.get()
        cfg = ctx.cfg
        s, p, x = ctx.saved_tensors[:3]
        adj = ctx.saved_tensors[3] if ctx.has_adj else None
        dev = s.device
        m, c = s.shape
        b, n = cfg.num_graphs, cfg.nmax
        g_ap = g_ap.contiguous() if g_ap is not None else torch.zeros(b, c, c, device=dev)
        g_xp = g_xp.contiguous() if g_xp is not None else torch.zeros(b, c, x.shape[1], device=dev)
        g_s = _mat(m, c, dev)
        be.gemm(_as3d(p, b, n), g_ap, _as3d(g_s, b, n), False, True, precision=_prec(be, 'pool_gs'))     # P gA'^T
        be.gemm(x.view(b, n, -1), g_xp, _as3d(g_s, b, n), False, True, beta=1.0,
                precision=_prec(be, 'pool_gs'))                                          # + X gX'^T
        g_p = _mat(m, c, dev)
        be.gemm(_as3d(s, b, n), g_ap, _as3d(g_p, b, n), False, False, precision=_prec(be, 'pool_gp'))    # S gA'
        g_adj = None
        if cfg.graph is not None:
            g = cfg.graph
            if _wide_ell_ok(g, g.ell_t, g_p, g_s):
                be.gather_wide_ell(g.ell_t, g.ell_t_val, None, g.ell_t_perm, g.ell_t_width, g.colptr, g.row,
                                   g.val_t, g_p, g.num_graphs, g.nmax, g_s, True)         # + A^T gP
            else:
                be.spmm(g.colptr, g.row, g.val_t, g_p, None, None, g_s, True)              # + A^T gP
        else:
            be.gemm(adj, _as3d(g_p, b, n), _as3d(g_s, b, n), True, False, beta=1.0)
            if ctx.want_adj:
                # (accumulating the three adjacency gradients of stage 2 into one buffer through beta = 1 epilogues was
                #  measured: the tensor-core kernel's row-per-thread epilogue reads the old tile uncoalesced, 1000 us per
                #  contraction instead of 209 us + 135 us for autograd's element-wise addition -- dropped)
                g_adj = torch.empty_like(adj)
                be.gemm(_as3d(g_p, b, n), _as3d(s, b, n), g_adj, False, True,
                        precision=_prec(be, 'pool_gadj'))                                # gP S^T
        g_x = _f32(m, x.shape[1], device=dev)
        be.gemm(_as3d(s, b, n), g_xp, g_x.view(b, n, -1), False, False, precision=_prec(be, 'pool_gx'))  # S gX'
        g_z = _mat(m, c, dev)                    # padded leading dimension: operand of the lin gradient contractions
        be.softmax_bwd(g_s, s, g_z)
        return None, g_z, g_x, g_adj, None


# ----------------------------------------------------------------------------- readout
class ColMaxFunction(Function):
    """Per-graph channel max over ALL ``Nmax`` rows, padded zero rows included (network.py:264)."""

    @staticmethod
    def forward(ctx, x, num_graphs, nmax):
        be = _backend.get()
        x = x.contiguous()
        out = _f32(num_graphs, x.shape[1], device=x.device)
        arg = torch.empty(num_graphs, x.shape[1], dtype=torch.int32, device=x.device)
        be.colmax_fwd(x, num_graphs, nmax, out, arg)
        ctx.arg, ctx.dims = arg, (num_graphs, nmax, x.shape[1])
        ctx.mark_non_differentiable(arg)
        return out, arg

    @staticmethod
    def backward(ctx, g, _garg):
        be = _backend.get()
        b, n, c = ctx.dims
        gx = torch.zeros(b * n, c, dtype=torch.float32, device=g.device)
        be.colmax_bwd(g.contiguous(), ctx.arg, b, n, gx)
        return gx, None, None


# ----------------------------------------------------------------------------- dense re-norm
class DenseRenormFunction(Function):
    """``_re_norm_adj(adj, p, mask)`` on a dense adjacency (network.py:183-191)."""

    @staticmethod
    def forward(ctx, adj, p, mask):
        be = _backend.get()
        adj = adj.contiguous()
        out = torch.empty_like(adj)
        rho = _f32(adj.shape[0] * adj.shape[1], device=adj.device)   # off-diagonal row sums + EPS
        be.dense_renorm_fwd(adj, p, mask, out, rho)
        ctx.p, ctx.mask = p, mask
        ctx.save_for_backward(out, rho)       # `out` is this node's own output: saved through autograd (no cycle)
        return out

    @staticmethod
    def backward(ctx, g):
        be = _backend.get()
        out, rho = ctx.saved_tensors
        g_adj = torch.empty_like(out)
        be.dense_renorm_bwd(g.contiguous(), out, rho, ctx.p, ctx.mask, g_adj)
        return g_adj, None, None


def dense_row_scale(adj):
    """``rsum = adj.sum(-1)`` and ``rs = 1 / clamp(rsum, min=1)`` of DenseSAGEConv on a dense
    adjacency.  Not an autograd node: its gradient w.r.t. ``adj`` is folded into
    ``GnnModuleFunction.backward`` (``add_rs_grad``)."""
    be = _backend.get()
    rows = adj.shape[0] * adj.shape[1]
    rs, rsum = _f32(rows, device=adj.device), _f32(rows, device=adj.device)
    be.row_inv_clamped_sum(adj.detach(), rs, rsum)
    return rs, rsum
