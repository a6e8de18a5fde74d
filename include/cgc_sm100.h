/*
 * cgc_sm100.h -- C ABI of libcgc_sm100.so: the B200 (sm_100a) graph-convolution hot path of
 * CGC-Net's SoftPoolingGcnEncoder (reference: /root/reference/model/network.py:57-291,
 * model/utils.py:3-37, and torch_geometric==1.2.1 DenseSAGEConv which the reference imports
 * at network.py:3).
 *
 * The reference has no FFI (pure Python on torch/PyG library kernels); this header is the new
 * seam.  Each entry point names the reference lines it replaces.
 *
 * Conventions
 *   - plain C: raw DEVICE pointers, explicit sizes and leading dimensions, cudaStream_t last
 *     (passed as void*).  No torch types.  fp32 row-major matrices, int32 CSR/CSC indices,
 *     int64 only for the reference's own COO / batch vectors.
 *   - "padded node layout": node i of graph b is row b * nmax + i of a [num_graphs * nmax, C]
 *     matrix (the reference's to_dense_batch layout, network.py:242).  A row mask is passed as
 *     (num_nodes[num_graphs], nmax): row r is real iff r % nmax < num_nodes[r / nmax];
 *     num_nodes == NULL means "no mask" (stages 2-3, network.py:272,281).
 *   - ownership: the caller allocates every input, output and workspace; the library never
 *     allocates, frees or keeps a pointer past the call.
 *   - functions only ENQUEUE work on `stream` of the calling thread's current device; no
 *     device synchronisation, no cudaSetDevice, re-entrant.
 *   - return value: 0 on success; a cudaError_t (> 0) or a negative argument-check code
 *     otherwise, with a thread-local message available from cgc_last_error().
 *     Never throws, never exits.
 *   - workspace: functions that take (work, work_bytes) need at least
 *     cgc_workspace_bytes() bytes, 256-byte aligned; one buffer may be shared by all calls
 *     on the same stream.  Its LAST 256 bytes hold self-resetting completion counters: they must be
 *     zero when the buffer is first handed to the library (the library leaves them zero).
 */
#ifndef CGC_SM100_H_
#define CGC_SM100_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CGC_API __attribute__((visibility("default")))
#else
#define CGC_API
#endif

#define CGC_OK 0
#define CGC_ERR_ARG (-1)        /* invalid argument (size, alignment, null pointer)      */
#define CGC_ERR_UNSUPPORTED (-2) /* shape outside what the kernels are instantiated for  */
#define CGC_ERR_WORKSPACE (-3)  /* workspace too small                                   */

/* activation after a conv (GNN_Module._activation, model/network.py:82-90), `act` arguments below */
#define CGC_ACT_RELU 0
#define CGC_ACT_ELU 1          /* alpha = 1 (nn.ELU default)           */
#define CGC_ACT_LEAKY_RELU 2   /* slope 0.01 (nn.LeakyReLU default)    */

CGC_API int cgc_version(void);
CGC_API const char* cgc_last_error(void);
/* Bytes of scratch that is enough for every entry point below at any supported shape. */
CGC_API int64_t cgc_workspace_bytes(void);
/* Number of CUDA kernels this library has launched so far in the process (bench.py's gpu_launches). */
CGC_API int64_t cgc_launch_count(void);

/* ------------------------------------------------------------------ graph construction
 * Replaces to_dense_adj (model/utils.py:3-37) + to_dense_batch (network.py:242) +
 * construct_mask (network.py:172-180).  No dense [B, Nmax, Nmax] tensor is built. */

/* num_nodes[g] = #{i : batch[i] == g}   (scatter_('add', ones, batch), model/utils.py:18-19) */
CGC_API int cgc_count_nodes(const int64_t* batch, int64_t n_total, int32_t num_graphs,
                    int32_t* num_nodes, void* stream);

/* out[batch[i] * nmax + (i - start[batch[i]]), :] = x[i, :], other rows 0   (to_dense_batch) */
CGC_API int cgc_pad_nodes(const float* x, const int64_t* batch, const int32_t* start, int64_t n_total,
                  int32_t feat, int32_t nmax, int32_t num_graphs, float* out, void* stream);

/* int32 elements of `work` needed by cgc_coo_to_csr / cgc_csr_transpose */
CGC_API int64_t cgc_coo_to_csr_workspace(int64_t edge_capacity, int64_t rows);

/* COO edge_index[2, n_edges] (int64, global node ids) -> CSR over the padded rows:
 * entry (r, c) with r = g*nmax + (src - start[g]), c = g*nmax + (dst - start[batch[dst]]),
 * g = batch[src]  (adj[g, src_local, dst_local] = 1, model/utils.py:29-33).  Duplicate edges
 * collapse; columns are sorted within a row.  add_self_loops != 0 forces the diagonal entry
 * of every real node (needed by the norm_adj edge weights, network.py:188).
 * rowptr[rows + 1], col[>= n_edges (+ n_total)].  The final nnz is rowptr[rows] (device). */
CGC_API int cgc_coo_to_csr(const int64_t* edge_index, int64_t n_edges, const int64_t* batch,
                   const int32_t* start, int64_t n_total, int32_t num_graphs, int32_t nmax,
                   int add_self_loops, int32_t* rowptr, int32_t* col, int32_t* work,
                   int64_t work_elems, void* stream);

/* rs[r] = 1 / max(deg(r), 1): the mean of DenseSAGEConv, adj.sum(-1).clamp(min=1) (PyG 1.2.1) */
CGC_API int cgc_row_inv_degree(const int32_t* rowptr, int64_t rows, float* rs, void* stream);

/* _re_norm_adj on the stored entries (network.py:183-191): val = p on the diagonal,
 * (1/(offdiag_degree + 1e-15)) * (1-p) elsewhere; rs[r] = 1 / max(sum_e val, 1). */
CGC_API int cgc_edge_renorm(const int32_t* rowptr, const int32_t* col, int64_t rows, double p,
                    float* val, float* rs, void* stream);

/* CSR -> CSC (rows sorted within a column), val (nullable) permuted into val_t. */
CGC_API int cgc_csr_transpose(const int32_t* rowptr, const int32_t* col, const float* val, int64_t rows,
                      int32_t* colptr, int32_t* row, float* val_t, int32_t* work,
                      int64_t work_elems, void* stream);

/* ------------------------------------------------------------------ cell-graph construction (loader side)
 * radius_graph(pos, r, batch, loop, max_num_neighbors = k) of torch_cluster 1.4.2 as called at
 * dataflow/data.py:246,255,297,348: for every centre c the <= k nearest nodes of the same graph with distance
 * < r (the centre itself first when loop != 0), nearest first, ties by lower index; distances in double on the
 * float positions (the CPU path is scipy cKDTree.query).  pos [n_total, 2]; start [num_graphs + 1] node offsets;
 * nbr [n_total, k] global node ids (-1 = none), deg [n_total] = valid entries per centre. */
CGC_API int cgc_radius_graph(const float* pos, const int32_t* start, int32_t num_graphs, int32_t nmax, double r,
                     int32_t k, int loop, int32_t* nbr, int32_t* deg, void* stream);
/* (nbr, deg, offsets = exclusive scan of deg) -> the reference's COO edge_index [2, num_edges] (int64):
 * edge_index[0] = neighbour, edge_index[1] = centre (flow = source_to_target), centre-major order. */
CGC_API int cgc_knn_to_coo(const int32_t* nbr, const int32_t* deg, const int64_t* offsets, int64_t n_total,
                   int32_t k, int64_t num_edges, int64_t* edge_index, void* stream);

/* Farthest-point node sampling = FarthestSampler.__call__ (common/utils.py:187-197) over the int16 distance
 * matrix of construct_feature_graph.py:17-24, d(i,j) = (int16) sqrt(|pos_i - pos_j|^2), evaluated on the fly:
 * out[out_start[g]] = first[g] (local id; the reference draws it with np.random.randint), then repeatedly the
 * first arg-max of the running min-distance.  start / out_start: [num_graphs + 1] offsets (nodes / samples);
 * out: local node ids.  The 'fuse' sampler (dataflow/data.py:210-219) takes 70 % of its nodes from this. */
CGC_API int cgc_farthest_point_sample(const float* pos, const int32_t* start, const int32_t* first,
                              const int32_t* out_start, int32_t num_graphs, int32_t nmax, int32_t* out,
                              void* stream);

/* ------------------------------------------------------------------ sparse neighbour gather
 * out[r,:] (+)= row_scale[r] * sum_{e in row r} val[e] * src_scale[col[e]] * x[col[e], :]
 * = torch.matmul(adj, x) / adj.sum(-1).clamp(min=1) of DenseSAGEConv.forward (forward: CSR +
 * row_scale = rs), its transpose (backward: CSC + src_scale = rs), and A*S of _diff_pool
 * (network.py:207).  val, src_scale, row_scale nullable (= 1). */
CGC_API int cgc_spmm(const int32_t* rowptr, const int32_t* col, const float* val, int64_t rows,
             const float* x, int64_t ldx, int32_t feat, const float* src_scale,
             const float* row_scale, float* out, int64_t ldo, int accumulate, void* stream);

/* The roofline kernel (BASELINE.json metric): mean-aggregation of one SAGE layer with the
 * per-graph feature tile staged in shared memory by bulk async copies (TMA, cp.async.bulk)
 * and neighbour rows gathered from shared memory.  out[r,:] = (1/max(deg,1)) * sum x[col,:]
 * (val == NULL) or rs-free weighted sum (val != NULL, row_scale nullable).
 * Requires: x, out contiguous [num_graphs*nmax, feat], feat % 4 == 0, column indices of graph g
 * inside [g*nmax, (g+1)*nmax), nmax*feat*4 + staging <= 200 KB; otherwise CGC_ERR_UNSUPPORTED
 * (callers then use cgc_spmm). */
CGC_API int cgc_sage_gather_mean_fwd(const int32_t* rowptr, const int32_t* col, const float* val,
                             const float* row_scale, const float* x, int32_t num_graphs,
                             int32_t nmax, int32_t feat, float* out, void* stream);

/* Row-padded CSR ("ELL16") of the stage-1 adjacency, the index layout of cgc_sage_gather_ell, cgc_sage_layer_ell_fwd
 * and cgc_gather_wide_ell.
 * Graph g owns nrow8 = round_up(nmax, 8) table rows (cgc_ell16_rows() = num_graphs * nrow8 in total); table row
 * g*nrow8 + 8*blk + p holds the entries of local row 8*blk + ((perm[g*nrow8/8 + blk] >> 3p) & 7):
 *   ell[row * width + t] = LOCAL column id (column - g*nmax, 12 bits of a uint16: nmax + 1 <= 4095) of its t-th
 *   entry, `nmax` in unused slots (and everywhere for rows beyond nmax); a row with more than `width` entries gets
 *   slot 0 = nmax + 1 and the gather walks the CSR for it.  Slot 0 additionally carries, in bits 12-14, the local row
 *   (0..7 within its block) the table row serves -- the same information as `perm`, readable without a second load.
 *   ell_val (nullable, with val) carries the entry weights, 0 in unused slots.
 * pair_rows != 0 chooses the slot order and the row-to-position map that minimise shared-memory bank conflicts
 * for 64-byte feature rows (slots go "even columns first" on even positions / "odd first" on odd ones; the
 * 8 rows of a block are paired smallest-with-largest even count); pair_rows == 0 keeps CSR order and the
 * identity map.  Same
 * adjacency as to_dense_adj (model/utils.py:29-33); only the summation order inside a row changes.
 * ell_scale[row] = the output scale of that table row: row_scale[r] when given (rs of cgc_edge_renorm /
 * cgc_row_inv_degree), else 1/max(deg,1) for a binary adjacency (val == NULL), else 1.
 * width is 8 or 16.  ell: cgc_ell16_rows()*width uint16; ell_scale: cgc_ell16_rows() floats;
 * perm: cgc_ell16_rows()/8 uint32. */
CGC_API int64_t cgc_ell16_rows(int32_t num_graphs, int32_t nmax);
CGC_API int cgc_csr_to_ell16(const int32_t* rowptr, const int32_t* col, const float* val,
                     const float* row_scale, int32_t num_graphs, int32_t nmax, int32_t width,
                     int pair_rows, uint16_t* ell, float* ell_val, float* ell_scale, uint32_t* perm,
                     void* stream);

/* The roofline kernel, fixed-pitch variant (BASELINE.json metric): same result as cgc_sage_gather_mean_fwd,
 * out[r,:] = scale[r] * sum_e val[e] * x[col[e],:] with scale / val as baked into the table by
 * cgc_csr_to_ell16 (binary adjacency: the DenseSAGEConv mean).  The per-graph feature tile is staged in
 * shared memory by bulk async copies (TMA); everything else a row needs (its whole index list = one 16-byte
 * load, weights, scale, position map) is addressed by the table row alone and issued at kernel start: no
 * rowptr -> column dependency, nothing but the tile goes through shared memory.  rowptr / col / val are read only for
 * rows marked as
 * overflowing `width`.  Requires (nmax + 2) * feat * 4 <= 226 KB, feat % 4 == 0, feat <= 32. */
CGC_API int cgc_sage_gather_ell_fwd(const uint16_t* ell, const float* ell_val, const float* ell_scale,
                            const uint32_t* perm, int32_t width, const int32_t* rowptr,
                            const int32_t* col, const float* val, const float* x, int32_t num_graphs,
                            int32_t nmax, int32_t feat, float* out, void* stream);

/* General form of the ELL16 gather: x and out may be column slices of wider matrices (leading dimensions ldx, ldo:
 * the layer-2/3 inputs of a GNN_Module live in the concat buffer, network.py:118; the CSC gathers of the backward
 * add into slices of the gradient buffer).  A strided x tile is staged by 2-D TMA boxes instead of 1-D bulk copies.
 * flags:
 *   CGC_GATHER_ACCUMULATE    out += result (default: out = result)
 *   CGC_GATHER_TABLES_READY  the index tables were NOT written by the kernel enqueued right before this call:
 *                            the kernel is launched with programmatic stream serialization and may then load its
 *                            indices / scales while the previous kernel drains (x is always read after it).
 *   CGC_GATHER_NO_PDL        launch without programmatic stream serialization. */
#define CGC_GATHER_ACCUMULATE 1
#define CGC_GATHER_TABLES_READY 2
#define CGC_GATHER_NO_PDL 4      /* plain stream-ordered launch (A/B measurements of the programmatic launch) */
#define CGC_GATHER_TWO_PER_SM 8  /* roofline configuration only (pitch 8, unweighted, no scale vector, 16 contiguous
                                    features): the 512-thread / 64-register variant of which two CTAs share an SM.
                                    It is what a programmatic launch of that configuration uses anyway (its CTAs
                                    prefetch their tile into L2 while the previous kernel runs); the flag forces it
                                    for a CGC_GATHER_NO_PDL launch too */
#define CGC_GATHER_ONE_PER_SM 16 /* ... and this one forces the 1024-thread one-CTA-per-SM kernel (A/B measurements) */
#define CGC_GATHER_NO_MULTICAST 32 /* two-per-SM kernel with 2 or 4 CTAs per graph: by default they form a thread-block
                                    cluster and share ONE multicast read of the tile (6.03 us per launch at the
                                    BASELINE shape); with this flag every CTA copies the whole tile itself (6.24 us) */
CGC_API int cgc_sage_gather_ell(const uint16_t* ell, const float* ell_val, const float* ell_scale,
                        const uint32_t* perm, int32_t width, const int32_t* rowptr, const int32_t* col,
                        const float* val, const float* x, int64_t ldx, int32_t num_graphs, int32_t nmax,
                        int32_t feat, float* out, int64_t ldo, int flags, void* stream);

/* One whole narrow SAGE layer up to the BatchNorm statistics in ONE kernel (SURVEY K3+K4(+K5 partial sums);
 * reference lines network.py:114-116 = DenseSAGEConv.forward + activation, and the batch statistics of bn_k):
 *   sa = scale * sum_e val[e] * x[col[e], :]          (the ELL16 gather above, tile staged by TMA)
 *   y  = act( l2normalize(sa @ W + b) * mask )         (cgc_sage_epilogue_fwd)
 *   mean / invstd / running stats of y over all `num_graphs * nmax` rows (cgc_bn_stats), when mean != NULL:
 *   per-CTA column partial sums in `work`, finalised in fixed order in double precision by whichever CTA
 *   finishes last (completion counter in the tail of `work`): no second launch.
 * sa [rows, cin] (contiguous) is written as well: the backward needs it (dW = sa^T dh).
 * Instantiated for (cin, cout) in {(16, 20), (20, 20)}: cgc_sage_layer_ell_supported(); other shapes take
 * cgc_sage_gather_ell + cgc_sage_layer_fwd_narrow.  flags as for cgc_sage_gather_ell (ACCUMULATE not allowed). */
CGC_API int cgc_sage_layer_ell_supported(int32_t cin, int32_t cout, int32_t nmax);
CGC_API int cgc_sage_layer_ell_fwd(const uint16_t* ell, const float* ell_val, const float* ell_scale,
                           const uint32_t* perm, int32_t width, const int32_t* rowptr, const int32_t* col,
                           const float* val, const float* x, int64_t ldx, int32_t num_graphs, int32_t nmax,
                           int32_t cin, const float* w, const float* b, const int32_t* num_nodes, int32_t cout,
                           int act, float* sa, float* y, int64_t ldy, float* running_mean, float* running_var,
                           float momentum, float eps, float* mean, float* invstd, void* work,
                           int64_t work_bytes, int flags, void* stream);

/* Wide gather over the same ELL16 table: out[r, :feat] (+)= scale[r] * sum_e val[e] * x[col[e], :feat] for wide
 * feature matrices -- A*S of _diff_pool (network.py:207, feat = 1140) and, with the table of the transposed
 * adjacency, its backward.  A work item is (graph, 16-column slice): the strided [nmax, 16] slice is staged in
 * shared memory once by 2-D TMA and gathered there (double-buffered, persistent CTAs), instead of fetching every
 * x row once per referencing edge through L2.  ell_scale nullable (= 1).  feat, ldx, ldo multiples of 4. */
CGC_API int cgc_gather_wide_ell(const uint16_t* ell, const float* ell_val, const float* ell_scale,
                        const uint32_t* perm, int32_t width, const int32_t* rowptr, const int32_t* col,
                        const float* val, const float* x, int64_t ldx, int32_t num_graphs, int32_t nmax,
                        int32_t feat, float* out, int64_t ldo, int accumulate, void* stream);

/* ------------------------------------------------------------------ SAGE layer epilogue
 * y = act( l2normalize(sa @ W + b) * mask )      [rows, cout]
 * DenseSAGEConv.forward lines matmul(out, weight) / + bias / F.normalize(p=2, eps=1e-12) /
 * * mask, then the in-place activation (network.py:82-90,114-116; act = CGC_ACT_*).
 * sa = mean-aggregated features [rows, cin]. */
CGC_API int cgc_sage_epilogue_fwd(const float* sa, int64_t ldsa, const float* w, const float* b,
                          const int32_t* num_nodes, int32_t nmax, int64_t rows, int32_t cin,
                          int32_t cout, int act, float* y, int64_t ldy, void* stream);

/* Wide counterpart (32 < cout <= 1280: the 20 -> 1140 assignment layer of GCN_pool_1, network.py:146-148): the same
 * epilogue with the BatchNorm column sums of y accumulated in the registers of the threads that own the columns,
 * one [2][cout] double partial per CTA in `work`, then the fixed-order finalize -- the separate 525 MB statistics
 * pass over y disappears. */
CGC_API int cgc_sage_layer_fwd_wide(const float* sa, int64_t ldsa, const float* w, const float* b,
                            const int32_t* num_nodes, int32_t nmax, int64_t rows, int32_t cin,
                            int32_t cout, int act, float* y, int64_t ldy, float* running_mean,
                            float* running_var, float momentum, float eps, float* mean, float* invstd,
                            void* work, int64_t work_bytes, void* stream);

/* Fused narrow layer forward (cout <= 32, cin <= 64, training mode): cgc_sage_epilogue_fwd AND cgc_bn_stats of
 * its output in one pass over the rows (thread per row, coalesced tile staging, per-CTA column-sum partials in
 * `work`, the same deterministic double-precision finalize): y as above; mean / invstd / running stats as
 * cgc_bn_stats(y). */
CGC_API int cgc_sage_layer_fwd_narrow(const float* sa, int64_t ldsa, const float* w, const float* b,
                              const int32_t* num_nodes, int32_t nmax, int64_t rows, int32_t cin,
                              int32_t cout, int act, float* y, int64_t ldy, float* running_mean,
                              float* running_var, float momentum, float eps, float* mean, float* invstd,
                              void* work, int64_t work_bytes, void* stream);

/* dh = d loss / d (sa @ W + b) given gy = d loss / d y.  dh may alias gy. */
CGC_API int cgc_sage_dh(const float* gy, int64_t ldgy, const float* sa, int64_t ldsa, const float* w,
                const float* b, const float* y, int64_t ldy, const int32_t* num_nodes,
                int32_t nmax, int64_t rows, int32_t cin, int32_t cout, int act, float* dh,
                int64_t lddh, void* stream);

/* cgc_sage_dh with the BatchNorm backward of the layer folded in (wide layers, 32 < cout <= 1280): gx is the
 * gradient w.r.t. the BatchNorm OUTPUT; d_gamma / d_beta are the column sums cgc_bn_bwd(..., gy = NULL, ...) has
 * produced.  dh = d loss / d (sa W + b) as above, with gy = gamma invstd (gx mask - d_beta / rows - xhat d_gamma / rows)
 * (training) or gamma invstd gx mask (eval) formed on the fly: the pass that wrote gy disappears. */
CGC_API int cgc_sage_dh_bn(const float* gx, int64_t ldgx, const float* sa, int64_t ldsa, const float* w,
                   const float* b, const float* y, int64_t ldy, const float* mean, const float* invstd,
                   const float* gamma, const float* d_gamma, const float* d_beta, int training,
                   const int32_t* num_nodes, int32_t nmax, int64_t rows, int32_t cin, int32_t cout, int act,
                   float* dh, int64_t lddh, void* stream);

/* Fused backward of a narrow layer epilogue (cout <= 32, cin <= 64) in one pass over the rows:
 * dsa = dh W^T [rows, cin], d_w = sa^T dh [cin, cout], d_b = sum_rows dh (nullable), with
 * dh = d loss / d (sa @ W + b) as in cgc_sage_dh (not materialised).  Deterministic (per-CTA partials in
 * `work`, fixed-order finalize). */
CGC_API int cgc_sage_layer_bwd_narrow(const float* gy, int64_t ldgy, const float* sa, int64_t ldsa,
                              const float* w, const float* b, const float* y, int64_t ldy,
                              const int32_t* num_nodes, int32_t nmax, int64_t rows, int32_t cin,
                              int32_t cout, int act, float* dsa, int64_t lddsa, float* d_w, float* d_b,
                              void* work, int64_t work_bytes, void* stream);

/* Element-wise activation y = act(x) and its gradient gx = gy * act'(y) (in terms of the output, as the
 * reference's in-place modules): the activation inside DenseGINConv's MLP and after it (network.py:96-99,114).
 * y may alias x; gx may alias gy. */
CGC_API int cgc_act_fwd(const float* x, int64_t ldx, int64_t rows, int32_t c, int act, float* y, int64_t ldy,
                void* stream);
CGC_API int cgc_act_bwd(const float* gy, int64_t ldgy, const float* y, int64_t ldy, int64_t rows, int32_t c,
                int act, float* gx, int64_t ldgx, void* stream);

/* Backward of sa = rs * agg (dense stages, where rs depends on the adjacency):
 * ga = rs * dsa (leading dimension ldga);  d_rs[r] += dot(dsa[r], sa[r]) / rs[r]. */
CGC_API int cgc_scale_rows_bwd(const float* dsa, const float* sa, const float* rs, int64_t rows,
                       int32_t c, float* ga, int64_t ldga, float* d_rs, void* stream);

/* ------------------------------------------------------------------ DenseJK (jumping knowledge, iff jk)
 * DenseJK.forward (model/network.py:36-52): x_t = feats[:, t*C:(t+1)*C] (t < layers); a bidirectional
 * LSTM(C -> H = C*layers/2, PyTorch gate order i,f,g,o, zero initial state) over the `layers` steps of every
 * node; score_t = att_w . [h_fwd_t, h_bwd_t] + att_b; alpha = softmax_t(score); out = sum_t alpha_t x_t.
 * Weight pointers are the nn.LSTM / nn.Linear parameters as PyTorch stores them (weight_ih_l0 [4H, C],
 * weight_hh_l0 [4H, H], bias_ih_l0, bias_hh_l0 [4H], the *_reverse set, att.weight [1, 2H], att.bias [1]).
 * tape (nullable = inference): cgc_jk_tape_floats() floats, sections gates [2][layers][5][H][rows] (i,f,g,o,c),
 * hs [2][layers][H][rows], alpha [layers][rows].  cgc_jk_supported() != 0 for the instantiated shapes
 * (channels = 20, layers = 3: the reference's hidden_dim); otherwise CGC_ERR_UNSUPPORTED. */
CGC_API int cgc_jk_supported(int32_t channels, int32_t layers);
CGC_API int64_t cgc_jk_tape_floats(int64_t rows, int32_t channels, int32_t layers);
CGC_API int cgc_jk_lstm_fwd(const float* feats, int64_t ld, int64_t rows, int32_t channels, int32_t layers,
                    const float* w_ih_f, const float* w_hh_f, const float* b_ih_f, const float* b_hh_f,
                    const float* w_ih_b, const float* w_hh_b, const float* b_ih_b, const float* b_hh_b,
                    const float* att_w, const float* att_b, float* out, int64_t ldo, float* tape,
                    void* stream);
/* Backward from the tape: g_feats [rows, layers*C]; dgates [2][layers][4H][rows] = d loss / d gate
 * pre-activations and dscore [layers][rows] = d loss / d score, from which the caller contracts the weight
 * gradients over the nodes (dW_ih = sum_t dgates_t x_t, dW_hh = sum_t dgates_t h_prev_t^T, db = row sums,
 * d att_w = sum_t hs_t dscore_t, d att_b = sum dscore) with cgc_gemm / cgc_colsum. */
CGC_API int cgc_jk_lstm_bwd(const float* g_out, int64_t ldg, const float* feats, int64_t ld, int64_t rows,
                    int32_t channels, int32_t layers, const float* w_ih_f, const float* w_hh_f,
                    const float* w_ih_b, const float* w_hh_b, const float* att_w, const float* tape,
                    float* g_feats, int64_t ldgf, float* dgates, float* dscore, void* stream);

/* ------------------------------------------------------------------ BatchNorm1d over ALL rows
 * (padded rows count, network.py:101-107).  Training: batch mean / biased variance, running
 * stats updated with momentum and the unbiased variance. */
CGC_API int cgc_bn_stats(const float* y, int64_t ldy, int64_t rows, int32_t c, float* running_mean,
                 float* running_var, float momentum, float eps, float* mean, float* invstd,
                 void* work, int64_t work_bytes, void* stream);
CGC_API int cgc_bn_eval_stats(const float* running_mean, const float* running_var, int32_t c, float eps,
                      float* mean, float* invstd, void* stream);
/* out = mask * ((y - mean) * invstd * gamma + beta); mean == NULL: out = mask * y */
CGC_API int cgc_bn_apply(const float* y, int64_t ldy, const float* mean, const float* invstd,
                 const float* gamma, const float* beta, const int32_t* num_nodes, int32_t nmax,
                 int64_t rows, int32_t c, float* out, int64_t ldo, void* stream);
/* d_gamma = sum g*xhat, d_beta = sum g (g = gx*mask);
 * gy = gamma*invstd*(g - d_beta/rows - xhat*d_gamma/rows)   (training)   or gamma*invstd*g */
/* gy may be NULL: only d_gamma / d_beta (the column sums) are produced */
CGC_API int cgc_bn_bwd(const float* gx, int64_t ldgx, const float* y, int64_t ldy, const float* mean,
               const float* invstd, const float* gamma, const int32_t* num_nodes, int32_t nmax,
               int64_t rows, int32_t c, int training, float* gy, int64_t ldgy, float* d_gamma,
               float* d_beta, void* work, int64_t work_bytes, void* stream);

/* ------------------------------------------------------------------ diff-pool pieces
 * s = softmax(z, -1) * mask (network.py:200-205); s_full (nullable) = softmax before masking
 * (collect_assign, :201-202); arg (nullable) = argmax_c z[r, c] = cluster index per node.
 * z and s_full are contiguous [rows, c]; s has leading dimension lds (>= c, padded for TMA). */
CGC_API int cgc_softmax_fwd(const float* z, int64_t rows, int32_t c, const int32_t* num_nodes,
                    int32_t nmax, float* s, int64_t lds, float* s_full, int32_t* arg, void* stream);
/* gz = s * (gs - sum_c gs * s);  gs and s share the leading dimension lds; gz has its own (ldgz >= c: a leading
 * dimension padded to 4 floats keeps the lin gradient contractions that consume gz on the tensor-core path) */
CGC_API int cgc_softmax_bwd(const float* gs, const float* s, int64_t lds, int64_t rows, int32_t c,
                    float* gz, int64_t ldgz, void* stream);

/* torch.max(x, dim=1) over all nmax rows of each graph (network.py:264,275,284) */
CGC_API int cgc_colmax_fwd(const float* x, int32_t num_graphs, int32_t nmax, int32_t c, float* out,
                   int32_t* arg, void* stream);
/* gx[g*nmax + arg[g,c], c] += g[g,c]  (gx zero-filled by the caller) */
CGC_API int cgc_colmax_bwd(const float* g, const int32_t* arg, int32_t num_graphs, int32_t nmax,
                   int32_t c, float* gx, void* stream);

/* ------------------------------------------------------------------ dense soft adjacency
 * rsum[r] = sum_j adj[r, j];  rs[r] = 1 / max(rsum[r], 1)   (DenseSAGEConv on stages 2-3) */
CGC_API int cgc_row_inv_clamped_sum(const float* adj, int64_t rows, int32_t n, float* rs, float* rsum,
                            void* stream);
/* d_adj[r, :] += (rsum[r] >= 1) ? -d_rs[r] / rsum[r]^2 : 0 */
CGC_API int cgc_add_rs_grad(float* d_adj, const float* rsum, const float* d_rs, int64_t rows, int32_t n,
                    void* stream);
/* _re_norm_adj(adj, p, mask) on a dense [num_graphs, n, n] adjacency (network.py:183-191);
 * rho[r] = off-diagonal row sum + 1e-15.  p is a double so that (float)(1 - p) equals the
 * reference's python-float arithmetic (1 - 0.4 -> 0.6f). */
CGC_API int cgc_dense_renorm_fwd(const float* adj, int32_t num_graphs, int32_t n, double p,
                         const int32_t* num_nodes, float* out, float* rho, void* stream);
CGC_API int cgc_dense_renorm_bwd(const float* g, const float* out, const float* rho, int32_t num_graphs,
                         int32_t n, double p, const int32_t* num_nodes, float* g_adj,
                         void* stream);

/* ------------------------------------------------------------------ dense contractions
 * C[i] = alpha * op(A[i]) @ op(B[i]) (+ bias[n]) (* row_scale[i*m + r]) (* mask(i*m + r))
 *        + beta * C[i],      i = 0 .. batch-1, op(X) = X or X^T, all fp32 row-major.
 * A is [m,k] (trans_a: stored [k,m]) with leading dimension lda and batch stride stride_a
 * (elements); likewise B [k,n] (trans_b: stored [n,k]) and C [m,n].
 * Serves lin of GNN_Module (network.py:121-124), S^T X and S^T (A S) of _diff_pool (:206-207),
 * torch.matmul(adj, x) of the dense stages, and all their gradients.
 * precision: 0 = fp32 FMA (exact fp32 products), 1 = tcgen05 3xTF32 (error ~1e-6 relative),
 *            -1 = library's choice (tcgen05 3xTF32 when the shape/alignment allows, else fp32 FMA),
 *            2 = like -1, but a contraction taken by the tensor-core path runs ONE TF32 pass on operands
 *                rounded to nearest (10-bit mantissa products, ~2e-4 relative per product, unbiased): for
 *                contractions whose consumers tolerate it (pass budget),
 *            3 = single TF32 pass on the tensor-core path or CGC_ERR_UNSUPPORTED (as 1 is to -1). */
CGC_API int cgc_gemm(int trans_a, int trans_b, int32_t m, int32_t n, int32_t k, float alpha,
             const float* a, int64_t lda, int64_t stride_a, const float* b, int64_t ldb,
             int64_t stride_b, float beta, float* c, int64_t ldc, int64_t stride_c,
             int32_t batch, const float* bias, const float* row_scale,
             const int32_t* num_nodes, int32_t nmax, int precision, void* work,
             int64_t work_bytes, void* stream);

/* out[c] = sum_r mask(r) * x[r, c] */
CGC_API int cgc_colsum(const float* x, int64_t ldx, int64_t rows, int32_t c, const int32_t* num_nodes,
               int32_t nmax, float* out, void* work, int64_t work_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CGC_SM100_H_ */
