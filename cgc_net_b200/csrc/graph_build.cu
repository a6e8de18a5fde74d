// Graph construction: COO edge list + batch vector -> CSR / CSC over the padded node rows.
// Replaces to_dense_adj (reference model/utils.py:3-37), to_dense_batch (network.py:242) and the
// stage-1 _re_norm_adj (network.py:183-191) without materialising a dense [B, Nmax, Nmax] tensor.
#include "common.cuh"

namespace cgc {

// ------------------------------------------------------------------ exclusive scan (int32)
constexpr int kScanThreads = 256;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ int block_exclusive_scan(int v, int* smem_warp, int& block_total) {
  // inclusive warp scan
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) smem_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = (lane < kScanThreads / 32) ? smem_warp[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < kScanThreads / 32) smem_warp[lane] = winc - w;   // exclusive warp offsets
    if (lane == kScanThreads / 32 - 1) smem_warp[32] = winc;    // block total
  }
  __syncthreads();
  block_total = smem_warp[32];
  int res = smem_warp[warp] + inc - v;
  __syncthreads();
  return res;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_local(const int* __restrict__ in, int* __restrict__ out,
                                                             int* __restrict__ sums, int64_t n) {
  __shared__ int smem_warp[33];
  int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int v[kScanItems];
  int tsum = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    v[i] = (base + i < n) ? in[base + i] : 0;
    tsum += v[i];
  }
  int total;
  int excl = block_exclusive_scan(tsum, smem_warp, total);
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < n) out[base + i] = excl;
    excl += v[i];
  }
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// single block: exclusive scan of the per-tile sums (sequential over chunks of kScanThreads)
__global__ void __launch_bounds__(kScanThreads) k_scan_sums(int* __restrict__ sums, int nblocks) {
  __shared__ int smem_warp[33];
  int carry = 0;
  for (int base = 0; base < nblocks; base += kScanThreads) {
    int i = base + threadIdx.x;
    int v = (i < nblocks) ? sums[i] : 0;
    int total;
    int excl = block_exclusive_scan(v, smem_warp, total);
    if (i < nblocks) sums[i] = excl + carry;
    carry += total;
  }
}

__global__ void __launch_bounds__(kScanThreads) k_scan_add(int* __restrict__ out, const int* __restrict__ sums,
                                                           int64_t n) {
  int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int add = sums[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kScanItems; ++i)
    if (base + i < n) out[base + i] += add;
}

// out[i] = sum_{j<i} in[j], i in [0, n).  `sums` needs ceil(n / kScanTile) ints.
static int exclusive_scan(const int* in, int* out, int* sums, int64_t n, cudaStream_t st) {
  int nblocks = (int)ceil_div(n, kScanTile);
  k_scan_local<<<nblocks, kScanThreads, 0, st>>>(in, out, sums, n);
  CGC_CHECK_LAUNCH("k_scan_local");
  if (nblocks > 1) {
    k_scan_sums<<<1, kScanThreads, 0, st>>>(sums, nblocks);
    CGC_CHECK_LAUNCH("k_scan_sums");
    k_scan_add<<<nblocks, kScanThreads, 0, st>>>(out, sums, n);
    CGC_CHECK_LAUNCH("k_scan_add");
  }
  return 0;
}

// ------------------------------------------------------------------ node bookkeeping
__global__ void k_count_nodes(const int64_t* __restrict__ batch, int64_t n, int num_graphs,
                              int* __restrict__ num_nodes) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    int64_t g = batch[i];
    if (g >= 0 && g < num_graphs) atomicAdd(num_nodes + g, 1);
  }
}

__global__ void k_pad_nodes(const float* __restrict__ x, const int64_t* __restrict__ batch,
                            const int* __restrict__ start, int64_t n, int feat, int nmax,
                            float* __restrict__ out) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * feat) return;
  int64_t i = idx / feat;
  int f = (int)(idx - i * feat);
  int64_t g = batch[i];
  int64_t row = g * nmax + (i - start[g]);
  out[row * feat + f] = x[idx];
}

// padded (row, col) of COO edge e:  adj[g, src - cum[g], dst - cum[batch[dst]]], g = batch[src]
__device__ __forceinline__ void edge_rc(const int64_t* __restrict__ ei, int64_t n_edges, int64_t e,
                                        const int64_t* __restrict__ batch, const int* __restrict__ start,
                                        int nmax, int64_t& r, int64_t& c) {
  int64_t src = ei[e], dst = ei[n_edges + e];
  int64_t g = batch[src];
  r = g * nmax + (src - start[g]);
  c = g * nmax + (dst - start[batch[dst]]);
}

__global__ void k_edge_count(const int64_t* __restrict__ ei, int64_t n_edges, const int64_t* __restrict__ batch,
                             const int* __restrict__ start, int64_t n_total, int nmax, int add_loops,
                             int* __restrict__ cnt) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n_edges) {
    int64_t r, c;
    edge_rc(ei, n_edges, e, batch, start, nmax, r, c);
    atomicAdd(cnt + r, 1);
  } else if (add_loops && e < n_edges + n_total) {
    int64_t i = e - n_edges;
    int64_t g = batch[i];
    atomicAdd(cnt + (g * nmax + (i - start[g])), 1);
  }
}

__global__ void k_edge_fill(const int64_t* __restrict__ ei, int64_t n_edges, const int64_t* __restrict__ batch,
                            const int* __restrict__ start, int64_t n_total, int nmax, int add_loops,
                            const int* __restrict__ rawptr, int* __restrict__ cursor, int* __restrict__ rawcol) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t r, c;
  if (e < n_edges) {
    edge_rc(ei, n_edges, e, batch, start, nmax, r, c);
  } else if (add_loops && e < n_edges + n_total) {
    int64_t i = e - n_edges;
    int64_t g = batch[i];
    r = c = g * nmax + (i - start[g]);
  } else {
    return;
  }
  int pos = atomicAdd(cursor + r, 1);
  rawcol[rawptr[r] + pos] = (int)c;
}

// one thread per row: insertion-sort the row's columns, collapse duplicates in place
__global__ void k_row_sort_unique(const int* __restrict__ rawptr, int* __restrict__ rawcol, int64_t rows,
                                  int* __restrict__ ucnt) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  int b = rawptr[r], e = rawptr[r + 1];
  for (int i = b + 1; i < e; ++i) {
    int v = rawcol[i];
    int j = i - 1;
    while (j >= b && rawcol[j] > v) {
      rawcol[j + 1] = rawcol[j];
      --j;
    }
    rawcol[j + 1] = v;
  }
  int u = 0;
  for (int i = b; i < e; ++i) {
    int v = rawcol[i];
    if (i == b || v != rawcol[b + u - 1]) rawcol[b + u++] = v;
  }
  ucnt[r] = u;
}

__global__ void k_compact(const int* __restrict__ rawptr, const int* __restrict__ rawcol,
                          const int* __restrict__ rowptr, int64_t rows, int* __restrict__ col) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  int src = rawptr[r], dst = rowptr[r], n = rowptr[r + 1] - dst;
  for (int i = 0; i < n; ++i) col[dst + i] = rawcol[src + i];
}

__global__ void k_row_inv_degree(const int* __restrict__ rowptr, int64_t rows, float* __restrict__ rs) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  int d = rowptr[r + 1] - rowptr[r];
  rs[r] = 1.0f / fmaxf((float)d, 1.0f);
}

// _re_norm_adj on the stored entries of one row (network.py:186-188) + DenseSAGEConv's clamp
__global__ void k_edge_renorm(const int* __restrict__ rowptr, const int* __restrict__ col, int64_t rows,
                              float p, float omp, float* __restrict__ val, float* __restrict__ rs) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  int b = rowptr[r], e = rowptr[r + 1];
  int off = 0;
  for (int i = b; i < e; ++i) off += (col[i] != (int)r);
  float w = (1.0f / ((float)off + kRenormEps)) * omp;
  float sum = 0.f;
  for (int i = b; i < e; ++i) {
    float v = (col[i] == (int)r) ? p : w;
    val[i] = v;
    sum += v;
  }
  rs[r] = 1.0f / fmaxf(sum, 1.0f);
}

// ------------------------------------------------------------------ transpose
__global__ void k_col_count(const int* __restrict__ col, const int* __restrict__ rowptr, int64_t rows,
                            int* __restrict__ ccnt) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  for (int i = rowptr[r]; i < rowptr[r + 1]; ++i) atomicAdd(ccnt + col[i], 1);
}

__global__ void k_col_fill(const int* __restrict__ rowptr, const int* __restrict__ col,
                           const float* __restrict__ val, int64_t rows, const int* __restrict__ colptr,
                           int* __restrict__ cursor, int* __restrict__ row_out, float* __restrict__ val_t) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  for (int i = rowptr[r]; i < rowptr[r + 1]; ++i) {
    int c = col[i];
    int pos = colptr[c] + atomicAdd(cursor + c, 1);
    row_out[pos] = (int)r;
    if (val != nullptr) val_t[pos] = val[i];
  }
}

__global__ void k_col_sort(const int* __restrict__ colptr, int64_t rows, int* __restrict__ row_out,
                           float* __restrict__ val_t) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= rows) return;
  int b = colptr[c], e = colptr[c + 1];
  for (int i = b + 1; i < e; ++i) {
    int v = row_out[i];
    float w = val_t ? val_t[i] : 0.f;
    int j = i - 1;
    while (j >= b && row_out[j] > v) {
      row_out[j + 1] = row_out[j];
      if (val_t) val_t[j + 1] = val_t[j];
      --j;
    }
    row_out[j + 1] = v;
    if (val_t) val_t[j + 1] = w;
  }
}

static inline int64_t scan_blocks(int64_t n) { return ceil_div(n, kScanTile) + 8; }

}  // namespace cgc

using namespace cgc;

extern "C" int cgc_count_nodes(const int64_t* batch, int64_t n_total, int32_t num_graphs, int32_t* num_nodes,
                               void* stream) {
  CGC_REQUIRE(batch && num_nodes && n_total >= 0 && num_graphs > 0, CGC_ERR_ARG, "cgc_count_nodes: bad args");
  if (n_total == 0) return 0;
  k_count_nodes<<<(unsigned)ceil_div(n_total, 256), 256, 0, as_stream(stream)>>>(batch, n_total, num_graphs,
                                                                                  num_nodes);
  CGC_CHECK_LAUNCH("k_count_nodes");
  return 0;
}

extern "C" int cgc_pad_nodes(const float* x, const int64_t* batch, const int32_t* start, int64_t n_total,
                             int32_t feat, int32_t nmax, int32_t num_graphs, float* out, void* stream) {
  CGC_REQUIRE(x && batch && start && out && feat > 0 && nmax > 0 && num_graphs > 0, CGC_ERR_ARG,
              "cgc_pad_nodes: bad args");
  cudaStream_t st = as_stream(stream);
  CGC_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)num_graphs * nmax * feat, st));
  if (n_total == 0) return 0;
  k_pad_nodes<<<(unsigned)ceil_div(n_total * feat, 256), 256, 0, st>>>(x, batch, start, n_total, feat, nmax, out);
  CGC_CHECK_LAUNCH("k_pad_nodes");
  return 0;
}

extern "C" int64_t cgc_coo_to_csr_workspace(int64_t edge_capacity, int64_t rows) {
  // cnt[rows+1] rawptr[rows+1] cursor[rows+1] ucnt[rows+1] rawcol[cap] sums[...]
  return 4 * (rows + 1) + edge_capacity + scan_blocks(rows + 1) + 64;
}

extern "C" int cgc_coo_to_csr(const int64_t* edge_index, int64_t n_edges, const int64_t* batch,
                              const int32_t* start, int64_t n_total, int32_t num_graphs, int32_t nmax,
                              int add_self_loops, int32_t* rowptr, int32_t* col, int32_t* work,
                              int64_t work_elems, void* stream) {
  CGC_REQUIRE(batch && start && rowptr && col && work && num_graphs > 0 && nmax > 0 && n_edges >= 0,
              CGC_ERR_ARG, "cgc_coo_to_csr: bad args");
  CGC_REQUIRE(n_edges == 0 || edge_index, CGC_ERR_ARG, "cgc_coo_to_csr: null edge_index");
  int64_t rows = (int64_t)num_graphs * nmax;
  int64_t cap = n_edges + (add_self_loops ? n_total : 0);
  CGC_REQUIRE(rows < (1ll << 31) - 2 && cap < (1ll << 31) - 2, CGC_ERR_UNSUPPORTED,
              "cgc_coo_to_csr: int32 index overflow");
  CGC_REQUIRE(work_elems >= cgc_coo_to_csr_workspace(cap, rows), CGC_ERR_WORKSPACE,
              "cgc_coo_to_csr: workspace too small");
  cudaStream_t st = as_stream(stream);
  int* cnt = work;
  int* rawptr = cnt + (rows + 1);
  int* cursor = rawptr + (rows + 1);
  int* ucnt = cursor + (rows + 1);
  int* rawcol = ucnt + (rows + 1);
  int* sums = rawcol + cap;
  CGC_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * (size_t)(rows + 1), st));
  CGC_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int) * (size_t)(rows + 1), st));
  CGC_CUDA(cudaMemsetAsync(ucnt, 0, sizeof(int) * (size_t)(rows + 1), st));
  if (cap > 0) {
    k_edge_count<<<(unsigned)ceil_div(cap, 256), 256, 0, st>>>(edge_index, n_edges, batch, start, n_total, nmax,
                                                               add_self_loops, cnt);
    CGC_CHECK_LAUNCH("k_edge_count");
  }
  int rc = exclusive_scan(cnt, rawptr, sums, rows + 1, st);
  if (rc) return rc;
  if (cap > 0) {
    k_edge_fill<<<(unsigned)ceil_div(cap, 256), 256, 0, st>>>(edge_index, n_edges, batch, start, n_total, nmax,
                                                              add_self_loops, rawptr, cursor, rawcol);
    CGC_CHECK_LAUNCH("k_edge_fill");
  }
  k_row_sort_unique<<<(unsigned)ceil_div(rows, 128), 128, 0, st>>>(rawptr, rawcol, rows, ucnt);
  CGC_CHECK_LAUNCH("k_row_sort_unique");
  rc = exclusive_scan(ucnt, rowptr, sums, rows + 1, st);
  if (rc) return rc;
  k_compact<<<(unsigned)ceil_div(rows, 128), 128, 0, st>>>(rawptr, rawcol, rowptr, rows, col);
  CGC_CHECK_LAUNCH("k_compact");
  return 0;
}

extern "C" int cgc_row_inv_degree(const int32_t* rowptr, int64_t rows, float* rs, void* stream) {
  CGC_REQUIRE(rowptr && rs && rows > 0, CGC_ERR_ARG, "cgc_row_inv_degree: bad args");
  k_row_inv_degree<<<(unsigned)ceil_div(rows, 256), 256, 0, as_stream(stream)>>>(rowptr, rows, rs);
  CGC_CHECK_LAUNCH("k_row_inv_degree");
  return 0;
}

extern "C" int cgc_edge_renorm(const int32_t* rowptr, const int32_t* col, int64_t rows, double p, float* val,
                               float* rs, void* stream) {
  CGC_REQUIRE(rowptr && col && val && rs && rows > 0, CGC_ERR_ARG, "cgc_edge_renorm: bad args");
  k_edge_renorm<<<(unsigned)ceil_div(rows, 128), 128, 0, as_stream(stream)>>>(rowptr, col, rows, (float)p,
                                                                              (float)(1.0 - p), val, rs);
  CGC_CHECK_LAUNCH("k_edge_renorm");
  return 0;
}

extern "C" int cgc_csr_transpose(const int32_t* rowptr, const int32_t* col, const float* val, int64_t rows,
                                 int32_t* colptr, int32_t* row, float* val_t, int32_t* work,
                                 int64_t work_elems, void* stream) {
  CGC_REQUIRE(rowptr && col && colptr && row && work && rows > 0, CGC_ERR_ARG, "cgc_csr_transpose: bad args");
  CGC_REQUIRE((val == nullptr) == (val_t == nullptr), CGC_ERR_ARG, "cgc_csr_transpose: val/val_t mismatch");
  CGC_REQUIRE(work_elems >= cgc_coo_to_csr_workspace(0, rows), CGC_ERR_WORKSPACE,
              "cgc_csr_transpose: workspace too small");
  cudaStream_t st = as_stream(stream);
  int* ccnt = work;
  int* cursor = ccnt + (rows + 1);
  int* sums = cursor + (rows + 1);
  CGC_CUDA(cudaMemsetAsync(ccnt, 0, sizeof(int) * (size_t)(2 * (rows + 1)), st));
  k_col_count<<<(unsigned)ceil_div(rows, 128), 128, 0, st>>>(col, rowptr, rows, ccnt);
  CGC_CHECK_LAUNCH("k_col_count");
  int rc = exclusive_scan(ccnt, colptr, sums, rows + 1, st);
  if (rc) return rc;
  k_col_fill<<<(unsigned)ceil_div(rows, 128), 128, 0, st>>>(rowptr, col, val, rows, colptr, cursor, row, val_t);
  CGC_CHECK_LAUNCH("k_col_fill");
  k_col_sort<<<(unsigned)ceil_div(rows, 128), 128, 0, st>>>(colptr, rows, row, val_t);
  CGC_CHECK_LAUNCH("k_col_sort");
  return 0;
}
