"""Per-operator CPU formulas for every C-ABI entry point.  TEST INFRASTRUCTURE.

``TorchBackend`` exposes the same method names and argument order as
``cgc_net_b200._backend.CudaBackend`` (one method per ``extern "C"`` function of
``include/cgc_sm100.h``) and computes each operator -- forward AND the hand-derived backward
 -- with plain torch ops on the tensors it is given, writing results
into the caller-allocated outputs exactly like the library does.

Used two ways, both from ``tests/`` only:
* injected with ``_backend.use_backend_for_testing`` so the autograd wiring in
  ``cgc_net_b200/ops.py`` can be checked against the dense oracle on a CPU-only host;
* as the op-level checker for each CUDA kernel in the ``-m gpu`` tests.

It is never reachable from the product path.
"""
import torch

L2_EPS = 1e-12          # F.normalize eps (torch 1.1 semantics)
RENORM_EPS = 1e-15      # model/network.py:8


def _valid_rows(mask, rows, device):
    """[rows] bool: row r is a real node (``r % nmax < num_nodes[r // nmax]``)."""
    if mask is None:
        return torch.ones(rows, dtype=torch.bool, device=device)
    r = torch.arange(rows, device=device)
    return (r % mask.nmax) < mask.num_nodes.to(device).long()[r // mask.nmax]


class TorchBackend(object):
    name = 'torch-op-reference'

    # ---------------------------------------------------------------- graph construction
    def count_nodes(self, batch, num_nodes):
        num_nodes.zero_()
        num_nodes.index_add_(0, batch, torch.ones_like(batch, dtype=torch.int32))

    def pad_nodes(self, x, batch, start, nmax, out):
        out.zero_()
        slot = torch.arange(batch.numel(), device=x.device) - start.long()[batch] + batch * nmax
        out[slot] = x

    def coo_to_csr_workspace(self, cap, rows):
        return 1

    def coo_to_csr(self, edge_index, batch, start, num_nodes, nmax, num_graphs, n_total,
                   add_self_loops, rowptr, col, work):
        # adj[batch[src], src - cum[batch[src]], dst - cum[batch[dst]]] = 1  (model/utils.py:29-33)
        src, dst = edge_index[0], edge_index[1]
        st = start.long()
        g = batch[src]
        r = g * nmax + (src - st[g])
        c = g * nmax + (dst - st[batch[dst]])
        if add_self_loops:
            node = torch.arange(n_total, device=src.device)
            d = batch * nmax + (node - st[batch])
            r, c = torch.cat([r, d]), torch.cat([c, d])
        rows = num_graphs * nmax
        key = torch.unique(r * rows + c)               # sorted, duplicates collapsed
        rr, cc = key // rows, key % rows
        counts = torch.zeros(rows, dtype=torch.long, device=src.device).index_add_(
            0, rr, torch.ones_like(rr))
        rowptr[0] = 0
        rowptr[1:] = counts.cumsum(0).to(torch.int32)
        col[:cc.numel()] = cc.to(torch.int32)

    def row_inv_degree(self, rowptr, rs):
        deg = (rowptr[1:] - rowptr[:-1]).float()
        rs.copy_(1.0 / deg.clamp(min=1))

    def edge_renorm(self, rowptr, col, p, val, rs):
        # _re_norm_adj (network.py:183-191) on the stored entries: diag := p,
        # off-diag := 1 / (offdiag_degree + EPS) * (1 - p);  rs = 1 / clamp(rowsum, 1)
        rows = rowptr.numel() - 1
        nnz = int(rowptr[-1])
        deg = (rowptr[1:] - rowptr[:-1]).long()
        row_of = torch.repeat_interleave(torch.arange(rows, device=col.device), deg)
        c = col[:nnz].long()
        is_diag = c == row_of
        off = torch.zeros(rows, device=col.device).index_add_(0, row_of, (~is_diag).float())
        v = torch.where(is_diag, torch.full_like(off[row_of], p),
                        (1.0 / (off[row_of] + RENORM_EPS)) * (1 - p))
        val[:nnz] = v
        rsum = torch.zeros(rows, device=col.device).index_add_(0, row_of, v)
        rs.copy_(1.0 / rsum.clamp(min=1))

    def csr_transpose(self, rowptr, col, val, colptr, row, val_t, work):
        rows = rowptr.numel() - 1
        nnz = int(rowptr[-1])
        deg = (rowptr[1:] - rowptr[:-1]).long()
        row_of = torch.repeat_interleave(torch.arange(rows, device=col.device), deg)
        c = col[:nnz].long()
        order = torch.argsort(c * rows + row_of)
        counts = torch.zeros(rows, dtype=torch.long, device=col.device).index_add_(
            0, c, torch.ones_like(c))
        colptr[0] = 0
        colptr[1:] = counts.cumsum(0).to(torch.int32)
        row[:nnz] = row_of[order].to(torch.int32)
        if val is not None:
            val_t[:nnz] = val[:nnz][order]

    # ---------------------------------------------------------------- cell-graph construction
    def radius_graph(self, pos, start, num_graphs, nmax, r, k, loop, nbr, deg):
        """torch_cluster 1.4.2 radius_graph on CPU = scipy cKDTree.query(k, distance_upper_bound=r) per graph
        (float64 distances on the float positions), nearest first, the centre itself first when ``loop``."""
        from scipy.spatial import cKDTree
        import numpy as np
        nbr.fill_(-1)
        st = start.tolist()
        for g in range(num_graphs):
            p = pos[st[g]:st[g + 1]].double().numpy()
            n = p.shape[0]
            if n == 0:
                continue
            kk = min(k + (0 if loop else 1), n)
            dist, idx = cKDTree(p).query(p, k=kk, distance_upper_bound=float(r))
            dist, idx = dist.reshape(n, kk), idx.reshape(n, kk)
            for c in range(n):
                cand = [(dist[c, s], int(idx[c, s])) for s in range(kk) if np.isfinite(dist[c, s])]
                if not loop:
                    cand = [t for t in cand if t[1] != c]
                cand = sorted(cand)[:k]
                for s, (_, j) in enumerate(cand):
                    nbr[st[g] + c, s] = st[g] + j
                deg[st[g] + c] = len(cand)

    def knn_to_coo(self, nbr, deg, offsets, k, num_edges, edge_index):
        valid = nbr >= 0
        centre = torch.arange(nbr.shape[0]).view(-1, 1).expand_as(nbr)
        edge_index[0, :num_edges] = nbr[valid].long()
        edge_index[1, :num_edges] = centre[valid]

    def farthest_point_sample(self, pos, start, first, out_start, nmax, out):
        """FarthestSampler.__call__ (common/utils.py:187-197) on the int16 distance matrix of
        construct_feature_graph.py:17-24 (euc_dist: float64 sqrt, astype(int16)), first index given."""
        import numpy as np
        st, ost, fi = start.tolist(), out_start.tolist(), first.tolist()
        for g in range(len(fi)):
            p = pos[st[g]:st[g + 1]].double().numpy()
            k = ost[g + 1] - ost[g]
            if p.shape[0] == 0 or k <= 0:
                continue
            ax = (p[:, 0, np.newaxis].T - p[:, 0, np.newaxis]) ** 2
            ay = (p[:, 1, np.newaxis].T - p[:, 1, np.newaxis]) ** 2
            arr = np.sqrt(ax + ay).astype(np.int16)
            far = np.zeros(k, dtype=np.int32)
            far[0] = fi[g]
            dist = arr[far[0]]
            for i in range(1, k):
                far[i] = np.argmax(dist)
                dist = np.minimum(dist, arr[far[i]])
            out[ost[g]:ost[g + 1]] = torch.from_numpy(far)

    # ---------------------------------------------------------------- sparse gather
    def spmm(self, rowptr, col, val, x, src_scale, row_scale, out, accumulate):
        rows = rowptr.numel() - 1
        nnz = int(rowptr[-1])
        deg = (rowptr[1:] - rowptr[:-1]).long()
        row_of = torch.repeat_interleave(torch.arange(rows, device=col.device), deg)
        c = col[:nnz].long()
        src = x[c]
        if src_scale is not None:
            src = src * src_scale[c][:, None]
        if val is not None:
            src = src * val[:nnz, None]
        res = torch.zeros(rows, x.shape[1], dtype=x.dtype, device=x.device).index_add_(0, row_of, src)
        if row_scale is not None:
            res = res * row_scale[:, None]
        if accumulate:
            out += res
        else:
            out.copy_(res)

    def sage_gather_mean_fwd(self, rowptr, col, val, row_scale, x, num_graphs, nmax, out):
        if val is None and row_scale is None:
            deg = (rowptr[1:] - rowptr[:-1]).float()
            row_scale = 1.0 / deg.clamp(min=1)
        self.spmm(rowptr, col, val, x, None, row_scale, out, False)

    def ell16_rows(self, num_graphs, nmax):
        return num_graphs * ((nmax + 7) // 8 * 8)

    def csr_to_ell16(self, rowptr, col, val, row_scale, num_graphs, nmax, width, pair_rows, ell, ell_val,
                     ell_scale, perm):
        """Row-padded CSR with uint16 local column ids (held in an int16 tensor), see include/cgc_sm100.h:
        slots parity-ordered by position and the 8 rows of a block paired smallest-with-largest even count
        (``pair_rows``); unused slots = nmax; overflow rows
        (degree > width) marked by slot 0 = nmax + 1; ``perm`` = 3 bits per position -> local row in the block."""
        nrow8 = (nmax + 7) // 8 * 8
        nb8 = nrow8 // 8
        rp, cl = rowptr.tolist(), col.tolist()
        vl = val.tolist() if val is not None else None
        table = torch.full((num_graphs * nrow8, width), nmax, dtype=torch.int64)
        wtab = torch.zeros(num_graphs * nrow8, width, dtype=torch.float32)
        stab = torch.ones(num_graphs * nrow8, dtype=torch.float32)
        pw_all = []
        n, pad_par = width, nmax & 1
        for blk in range(num_graphs * nb8):
            g, lb = divmod(blk, nb8)
            base = g * nmax
            deg, key = [-1] * 8, [n // 2] * 8
            for q in range(8):
                lr = lb * 8 + q
                if lr >= nmax:
                    continue
                b, e = rp[base + lr], rp[base + lr + 1]
                deg[q] = e - b
                if deg[q] > width:
                    continue
                if deg[q] == 0:
                    key[q] = n if pad_par == 0 else 0
                    continue
                ev = sum(1 for i in range(b, e) if ((cl[i] - base) & 1) == 0)
                key[q] = ev + ((width - deg[q]) if pad_par == 0 else 0)
            order = sorted(range(8), key=lambda q: key[q]) if pair_rows else list(range(8))   # stable
            pw = 0
            for p in range(8):
                q = (order[7 - (p >> 1)] if (p & 1) else order[p >> 1]) if pair_rows else p
                pw |= q << (3 * p)
                trow = blk * 8 + p
                if deg[q] >= 0:
                    if row_scale is not None:
                        stab[trow] = row_scale[base + lb * 8 + q]
                    elif val is None:
                        stab[trow] = 1.0 / torch.tensor(float(max(deg[q], 1)), dtype=torch.float32)
                if deg[q] > width:
                    table[trow, 0] = nmax + 1
                elif deg[q] > 0:
                    b = rp[base + lb * 8 + q]
                    e = b + deg[q]
                    ent = list(range(b, e))
                    if pair_rows:
                        want = p & 1
                        ent = [i for i in ent if ((cl[i] - base) & 1) == want] + \
                              [i for i in ent if ((cl[i] - base) & 1) != want]
                    for t, i in enumerate(ent):
                        table[trow, t] = cl[i] - base
                        if vl is not None:
                            wtab[trow, t] = vl[i]
            pw_all.append(pw)
            for p in range(8):                      # slot 0 also carries the served local row (bits 12-14)
                table[blk * 8 + p, 0] |= ((pw >> (3 * p)) & 7) << 12
        table = torch.where(table >= 32768, table - 65536, table).to(torch.int16)
        ell.view(-1)[:table.numel()].copy_(table.view(-1))
        if ell_val is not None:
            ell_val.view(-1)[:wtab.numel()].copy_(wtab.view(-1))
        ell_scale.view(-1)[:stab.numel()].copy_(stab)
        perm.view(-1)[:len(pw_all)].copy_(torch.tensor(pw_all, dtype=torch.int64).to(torch.int32))

    def sage_gather_ell_fwd(self, ell, ell_val, ell_scale, perm, width, rowptr, col, val, x, num_graphs, nmax,
                            out):
        """Same sums as sage_gather_mean_fwd, taken over the ELL16 table (overflow rows over the CSR)."""
        nrow8 = (nmax + 7) // 8 * 8
        trows = num_graphs * nrow8
        idx = ell.view(-1)[:trows * width].view(trows, width).long() & 0xFFFF
        served = (idx[:, 0] >> 12) & 7
        idx[:, 0] &= 0xFFF
        pw = perm.view(-1)[:trows // 8].long() & 0xFFFFFF
        pos = torch.arange(trows)
        lrow = (pos % nrow8) // 8 * 8 + ((pw[pos // 8] >> (3 * (pos % 8))) & 7)             # local row per table row
        assert bool((served == ((pw[pos // 8] >> (3 * (pos % 8))) & 7)).all()), 'slot-0 row tag disagrees with perm'
        gid = pos // nrow8
        real = lrow < nmax
        xz = torch.cat([x.view(num_graphs, nmax, -1),
                        torch.zeros(num_graphs, 2, x.shape[1], dtype=x.dtype)], dim=1)     # zero rows at nmax, nmax+1
        src = xz[gid.view(-1, 1), idx.clamp(max=nmax + 1)]                                # [trows, width, F]
        if ell_val is not None:
            res = (src * ell_val.view(-1)[:trows * width].view(trows, width, 1)).sum(1)
        else:
            res = src.sum(1)
        grow = gid * nmax + lrow.clamp(max=nmax - 1)
        if ell_scale is None:                   # unweighted table: the mean 1 / max(deg, 1), deg from the row itself
            deg = (rowptr[1:] - rowptr[:-1]).float()[grow]
            scale = 1.0 / torch.where(real, deg, torch.ones_like(deg)).clamp(min=1)
        else:
            scale = ell_scale.view(-1)[:trows]
        res = res * scale[:, None]
        out[grow[real]] = res[real]
        over = real & (idx[:, 0] == nmax + 1)
        if bool(over.any()):                    # rows above the pitch: plain CSR sum, same table scale
            full = torch.zeros(num_graphs * nmax, x.shape[1], dtype=x.dtype)
            self.spmm(rowptr, col, val, x, None, None, full, False)
            out[grow[over]] = full[grow[over]] * scale[over][:, None]

    def sage_gather_ell(self, ell, ell_val, ell_scale, perm, width, rowptr, col, val, x, num_graphs, nmax, out,
                        accumulate=False, tables_ready=False, pdl=True, two_per_sm=None, multicast=True):
        """cgc_sage_gather_ell: strided x / out, optional +=."""
        res = torch.zeros(num_graphs * nmax, x.shape[1], dtype=x.dtype)
        self.sage_gather_ell_fwd(ell, ell_val, ell_scale, perm, width, rowptr, col, val, x.contiguous(), num_graphs,
                                 nmax, res)
        if accumulate:
            out += res
        else:
            out.copy_(res)

    @staticmethod
    def sage_layer_ell_supported(cin, cout, nmax):
        return cout == 20 and cin in (16, 20) and 0 < nmax < 0xFFF and ((nmax + 255) // 8 * 8) * cin * 4 <= 227 * 1024 - 2048

    def sage_layer_ell_fwd(self, ell, ell_val, ell_scale, perm, width, rowptr, col, val, x, num_graphs, nmax, w, b,
                           mask, act, sa, y, rmean, rvar, momentum, eps, mean, invstd, tables_ready=False, pdl=True):
        """cgc_sage_layer_ell_fwd = the ELL16 gather, then the SAGE epilogue, then the batch statistics."""
        self.sage_gather_ell(ell, ell_val, ell_scale, perm, width, rowptr, col, val, x, num_graphs, nmax, sa)
        self.sage_epilogue_fwd(sa, w, b, mask, y, act)
        if mean is not None:
            self.bn_stats(y, rmean, rvar, momentum, eps, mean, invstd)

    def gather_wide_ell(self, ell, ell_val, ell_scale, perm, width, rowptr, col, val, x, num_graphs, nmax, out,
                        accumulate):
        """Same sums as sage_gather_ell_fwd on a wide (possibly strided) matrix, optional scale, optional +=."""
        trows = num_graphs * ((nmax + 7) // 8 * 8)
        scale = ell_scale if ell_scale is not None else torch.ones(trows)
        res = torch.zeros(num_graphs * nmax, x.shape[1], dtype=x.dtype)
        self.sage_gather_ell_fwd(ell, ell_val, scale, perm, width, rowptr, col, val, x.contiguous(), num_graphs, nmax,
                                 res)
        if accumulate:
            out += res
        else:
            out.copy_(res)

    @staticmethod
    def _act(act, v):
        if act == 0:
            return v.clamp(min=0)
        if act == 1:
            return torch.where(v > 0, v, torch.expm1(v))
        return torch.where(v > 0, v, 0.01 * v)

    @staticmethod
    def _act_grad(act, y):
        if act == 0:
            return (y > 0).float()
        if act == 1:
            return torch.where(y > 0, torch.ones_like(y), y + 1)
        return torch.where(y > 0, torch.ones_like(y), torch.full_like(y, 0.01))

    def act_fwd(self, x, act, y):
        y.copy_(self._act(act, x))

    def act_bwd(self, gy, y, act, gx):
        gx.copy_(gy * self._act_grad(act, y))

    def _epilogue(self, sa, w, b, mask):
        h = sa @ w
        if b is not None:
            h = h + b
        nrm = h.norm(dim=1, keepdim=True)
        inv = 1.0 / nrm.clamp(min=L2_EPS)
        valid = _valid_rows(mask, sa.shape[0], sa.device)[:, None].float()
        return h, nrm, inv, valid

    def sage_epilogue_fwd(self, sa, w, b, mask, y, act=0):
        h, nrm, inv, valid = self._epilogue(sa, w, b, mask)
        y.copy_(self._act(act, h * inv * valid))

    def sage_layer_fwd_narrow(self, sa, w, b, mask, y, act, rmean, rvar, momentum, eps, mean, invstd):
        self.sage_epilogue_fwd(sa, w, b, mask, y, act)
        self.bn_stats(y, rmean, rvar, momentum, eps, mean, invstd)

    sage_layer_fwd_wide = sage_layer_fwd_narrow      # same formulas; the library splits them by output width

    def sage_dh(self, gy, sa, w, b, y, mask, dh, act=0):
        h, nrm, inv, valid = self._epilogue(sa, w, b, mask)
        hn = h * inv
        gm = gy * self._act_grad(act, y) * valid         # activation gradient (from the output) and row mask
        dot = (gm * hn).sum(1, keepdim=True)
        # y = h / max(|h|, eps): d/dh = (g - hn (hn . g)) / |h|   when |h| >= eps, g / eps otherwise
        dh.copy_(torch.where(nrm >= L2_EPS, (gm - hn * dot) * inv, gm * inv))

    def sage_dh_bn(self, gx, sa, w, b, y, mean, invstd, gamma, d_gamma, d_beta, training, mask, dh, act=0):
        """cgc_sage_dh_bn: the BatchNorm backward apply step (from the given column sums), then sage_dh."""
        m = y.shape[0]
        valid = _valid_rows(mask, m, y.device)[:, None].float()
        g = gx * valid
        if training:
            gy = gamma * invstd * (g - d_beta / m - (y - mean) * invstd * d_gamma / m)
        else:
            gy = gamma * invstd * g
        self.sage_dh(gy, sa, w, b, y, mask, dh, act)

    def sage_layer_bwd_narrow(self, gy, sa, w, b, y, mask, dsa, d_w, d_b, act=0):
        dh = torch.empty_like(y)
        self.sage_dh(gy, sa, w, b, y, mask, dh, act)
        dsa.copy_(dh @ w.t())
        d_w.copy_(sa.t() @ dh)
        if d_b is not None:
            d_b.copy_(dh.sum(0))

    def scale_rows_bwd(self, dsa, sa, rs, ga, d_rs):
        ga.copy_(dsa * rs[:, None])
        d_rs += (dsa * sa).sum(1) / rs

    # ---------------------------------------------------------------- DenseJK (bi-LSTM attention over layers)
    @staticmethod
    def jk_supported(channels, layers):
        return (channels * layers) % 2 == 0

    @staticmethod
    def jk_tape_floats(rows, channels, layers):
        hid = channels * layers // 2
        return rows * (2 * layers * 5 * hid + 2 * layers * hid + layers)

    @staticmethod
    def _jk_tape_views(tape, rows, channels, layers):
        hid = channels * layers // 2
        n_g = 2 * layers * 5 * hid * rows
        n_h = 2 * layers * hid * rows
        gates = tape[:n_g].view(2, layers, 5, hid, rows)                  # i, f, g, o, c  (node fastest)
        hs = tape[n_g:n_g + n_h].view(2, layers, hid, rows)
        alpha = tape[n_g + n_h:n_g + n_h + layers * rows].view(layers, rows)
        return gates, hs, alpha

    def jk_lstm_fwd(self, feats, channels, layers, lstm_params, att_w, att_b, out, tape):
        """DenseJK.forward (model/network.py:36-52): bi-LSTM(C -> H = C*L/2) over the L per-layer features of every
        node (PyTorch gate order i, f, g, o), Linear(2H -> 1) score, softmax over layers, weighted sum."""
        rows = feats.shape[0]
        hid = channels * layers // 2
        xs = [feats[:, t * channels:(t + 1) * channels] for t in range(layers)]
        if tape is not None:
            gates, hs, alpha_t = self._jk_tape_views(tape, rows, channels, layers)
        hall = torch.zeros(2, layers, rows, hid, dtype=feats.dtype)
        for d in range(2):
            w_ih, w_hh, b_ih, b_hh = lstm_params[4 * d:4 * d + 4]
            h = torch.zeros(rows, hid, dtype=feats.dtype)
            c = torch.zeros(rows, hid, dtype=feats.dtype)
            order = range(layers) if d == 0 else range(layers - 1, -1, -1)
            for t in order:
                a = xs[t] @ w_ih.t() + h @ w_hh.t() + (b_ih + b_hh)
                i, f = torch.sigmoid(a[:, :hid]), torch.sigmoid(a[:, hid:2 * hid])
                g, o = torch.tanh(a[:, 2 * hid:3 * hid]), torch.sigmoid(a[:, 3 * hid:])
                c = f * c + i * g
                h = o * torch.tanh(c)
                hall[d, t] = h
                if tape is not None:
                    for q, v in enumerate((i, f, g, o, c)):
                        gates[d, t, q].copy_(v.t())
                    hs[d, t].copy_(h.t())
        score = torch.stack([torch.cat([hall[0, t], hall[1, t]], 1) @ att_w.view(-1) + att_b.view(())
                             for t in range(layers)], 1)                       # [rows, L]
        alpha = torch.softmax(score, dim=1)
        if tape is not None:
            alpha_t.copy_(alpha.t())
        out.copy_(sum(alpha[:, t:t + 1] * xs[t] for t in range(layers)))

    def jk_lstm_bwd(self, g_out, feats, channels, layers, lstm_params, att_w, tape, g_feats, dgates, dscore):
        """Backward of jk_lstm_fwd from the tape: g_feats [rows, L*C]; dgates [2, L, 4H, rows] = d loss / d gate
        pre-activations (weight / bias gradients are contractions of these, done by the caller); dscore [L, rows]."""
        rows = feats.shape[0]
        hid = channels * layers // 2
        gates, hs, alpha = self._jk_tape_views(tape, rows, channels, layers)
        xs = [feats[:, t * channels:(t + 1) * channels] for t in range(layers)]
        d_alpha = torch.stack([(g_out * xs[t]).sum(1) for t in range(layers)], 0)          # [L, rows]
        dot = (alpha * d_alpha).sum(0, keepdim=True)
        ds = alpha * (d_alpha - dot)
        dscore.copy_(ds)
        gx = [alpha[t][:, None] * g_out for t in range(layers)]
        aw = att_w.view(-1)
        for d in range(2):
            w_ih, w_hh = lstm_params[4 * d], lstm_params[4 * d + 1]
            order = list(range(layers)) if d == 0 else list(range(layers - 1, -1, -1))
            dh_next = torch.zeros(rows, hid, dtype=feats.dtype)
            dc_next = torch.zeros(rows, hid, dtype=feats.dtype)
            for pos in range(layers - 1, -1, -1):
                t = order[pos]
                i, f, g, o, c = [gates[d, t, q].t() for q in range(5)]
                c_prev = gates[d, order[pos - 1], 4].t() if pos > 0 else torch.zeros_like(c)
                dh = ds[t][:, None] * aw[d * hid:(d + 1) * hid][None, :] + dh_next
                tc = torch.tanh(c)
                dc = dh * o * (1 - tc * tc) + dc_next
                da = torch.cat([dc * g * i * (1 - i), dc * c_prev * f * (1 - f), dc * i * (1 - g * g),
                                dh * tc * o * (1 - o)], 1)                                # [rows, 4H]
                dgates[d, t].copy_(da.t())
                dc_next = dc * f
                dh_next = da @ w_hh
                gx[t] = gx[t] + da @ w_ih
        g_feats.copy_(torch.cat(gx, 1))

    # ---------------------------------------------------------------- batch norm
    def bn_stats(self, y, rmean, rvar, momentum, eps, mean, invstd):
        m = y.shape[0]
        mu = y.mean(0)
        var = y.var(0, unbiased=False)
        mean.copy_(mu)
        invstd.copy_(torch.rsqrt(var + eps))
        rmean.mul_(1 - momentum).add_(momentum * mu)
        rvar.mul_(1 - momentum).add_(momentum * var * (m / max(m - 1, 1)))

    def bn_eval_stats(self, rmean, rvar, eps, mean, invstd):
        mean.copy_(rmean)
        invstd.copy_(torch.rsqrt(rvar + eps))

    def bn_apply(self, y, mean, invstd, gamma, beta, mask, out):
        valid = _valid_rows(mask, y.shape[0], y.device)[:, None].float()
        if mean is None:
            out.copy_(y * valid)
        else:
            out.copy_(((y - mean) * invstd * gamma + beta) * valid)

    def bn_bwd(self, gx, y, mean, invstd, gamma, mask, training, gy, d_gamma, d_beta):
        m = y.shape[0]
        valid = _valid_rows(mask, m, y.device)[:, None].float()
        g = gx * valid
        xhat = (y - mean) * invstd
        dg, db = (g * xhat).sum(0), g.sum(0)
        d_gamma.copy_(dg)
        d_beta.copy_(db)
        if gy is None:                  # column sums only (the consumer folds the apply step: sage_dh_bn)
            return
        if training:
            gy.copy_(gamma * invstd * (g - db / m - xhat * dg / m))
        else:
            gy.copy_(gamma * invstd * g)

    # ---------------------------------------------------------------- softmax / readout
    def softmax_fwd(self, z, mask, s, s_full, arg):
        p = torch.softmax(z, dim=1)
        valid = _valid_rows(mask, z.shape[0], z.device)[:, None].float()
        s.copy_(p * valid)
        if s_full is not None:
            s_full.copy_(p)
        if arg is not None:
            arg.copy_(z.argmax(dim=1).to(torch.int32))

    def softmax_bwd(self, gs, s, gz):
        gz.copy_(s * (gs - (gs * s).sum(1, keepdim=True)))

    def colmax_fwd(self, x, b, n, out, arg):
        v, i = x.view(b, n, -1).max(dim=1)
        out.copy_(v)
        arg.copy_(i.to(torch.int32))

    def colmax_bwd(self, g, arg, b, n, gx):
        gx.view(b, n, -1).scatter_add_(1, arg.long().unsqueeze(1), g.unsqueeze(1))

    # ---------------------------------------------------------------- dense adjacency helpers
    def row_inv_clamped_sum(self, adj, rs, rsum):
        r = adj.sum(-1).reshape(-1)
        rsum.copy_(r)
        rs.copy_(1.0 / r.clamp(min=1))

    def add_rs_grad(self, d_adj, rsum, d_rs):
        v = torch.where(rsum >= 1, -d_rs / (rsum * rsum), torch.zeros_like(rsum))
        d_adj += v.view(d_adj.shape[0], d_adj.shape[1], 1)

    def dense_renorm_fwd(self, adj, p, mask, out, rho):
        b, n, _ = adj.shape
        eye = torch.eye(n, dtype=torch.bool, device=adj.device)
        a0 = adj.masked_fill(eye, 0.0)
        r = a0.sum(-1, keepdim=True) + RENORM_EPS
        o = a0 / r * (1 - p)
        o = o.masked_fill(eye, p)
        valid = _valid_rows(mask, b * n, adj.device).view(b, n, 1).float()
        out.copy_(o * valid)
        rho.copy_(r.reshape(-1))

    def dense_renorm_bwd(self, g, out, rho, p, mask, g_adj):
        # N_ij = (1-p) A_ij / rho_i (i != j); diag constant.  dA_ij = (1-p)/rho_i (g_ij - sum_k g_ik N_ik/(1-p))
        b, n, _ = g.shape
        eye = torch.eye(n, dtype=torch.bool, device=g.device)
        valid = _valid_rows(mask, b * n, g.device).view(b, n, 1).float()
        gm = (g * valid).masked_fill(eye, 0.0)
        nn_ = out.masked_fill(eye, 0.0)              # masked rows are already 0
        # for masked-out rows `out` is 0 but the un-masked N is needed only where gm != 0 (valid rows)
        dot = (gm * nn_).sum(-1, keepdim=True) / (1 - p)
        ga = (1 - p) / rho.view(b, n, 1) * (gm - dot)
        g_adj.copy_(ga.masked_fill(eye, 0.0))

    # ---------------------------------------------------------------- dense contractions
    def gemm(self, a, b, c, ta, tb, alpha=1.0, beta=0.0, bias=None, row_scale=None, rowmask=None,
             precision=-1):
        aa = a.transpose(-1, -2) if ta else a
        bb = b.transpose(-1, -2) if tb else b
        r = alpha * (aa @ bb)
        if bias is not None:
            r = r + bias
        if row_scale is not None:
            r = r * row_scale.view(*r.shape[:-1], 1)
        if rowmask is not None:
            rows = r.shape[-2] if r.dim() == 2 else r.shape[0] * r.shape[1]
            valid = _valid_rows(rowmask, rows, r.device).float()
            r = r * valid.view(*r.shape[:-1], 1)
        if beta != 0.0:
            c.mul_(beta).add_(r)
        else:
            c.copy_(r)

    def colsum(self, x, mask, out):
        valid = _valid_rows(mask, x.shape[0], x.device)[:, None].float()
        out.copy_((x * valid).sum(0))
