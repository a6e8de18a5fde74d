// Cell-graph construction on the GPU: radius_graph(pos, r, batch, loop=True, max_num_neighbors=k) of
// torch_cluster 1.4.2 as the reference calls it (dataflow/data.py:246,255,297,348): for every centre node the
// <= k nearest nodes of the same graph within distance r (the centre itself first, at distance 0), returned as
// COO pairs (edge_index[0] = neighbour, edge_index[1] = centre).  The CPU implementation behind the reference is
// scipy's cKDTree.query(k, distance_upper_bound=r); here one thread owns one centre, scans its graph's positions
// from shared memory and keeps the k best (distance, index) pairs in registers.  Distances are evaluated in
// double precision on the float positions (as cKDTree does), ties broken by the lower node index.
#include "common.cuh"

namespace cgc {

constexpr int kKnnThreads = 128;

template <int K>
__global__ void __launch_bounds__(kKnnThreads)
k_radius_graph(const float* __restrict__ pos, const int* __restrict__ start, double r2, int k, int loop,
               int* __restrict__ nbr, int* __restrict__ deg) {
  extern __shared__ float2 s_pos[];
  const int g = blockIdx.y;
  const int n0 = __ldg(start + g), n = __ldg(start + g + 1) - n0;
  if ((int)(blockIdx.x * kKnnThreads) >= n) return;
  for (int i = threadIdx.x; i < n; i += kKnnThreads)
    s_pos[i] = make_float2(__ldg(pos + 2 * (int64_t)(n0 + i)), __ldg(pos + 2 * (int64_t)(n0 + i) + 1));
  __syncthreads();
  const int c = blockIdx.x * kKnnThreads + threadIdx.x;
  if (c >= n) return;
  const double cx = s_pos[c].x, cy = s_pos[c].y;
  double bd[K];
  int bi[K];
#pragma unroll
  for (int s = 0; s < K; ++s) { bd[s] = 1e300; bi[s] = -1; }
  for (int j = 0; j < n; ++j) {
    if (!loop && j == c) continue;
    const double dx = (double)s_pos[j].x - cx, dy = (double)s_pos[j].y - cy;
    double d = dx * dx + dy * dy;
    if (d >= r2 && j != c) continue;                       // cKDTree: strictly inside the upper bound
    if (!(d < bd[K - 1])) continue;                        // not better than the current k-th (ties: keep lower index)
    int id = j;
#pragma unroll
    for (int s = 0; s < K; ++s) {                          // insertion into the sorted list
      const bool better = d < bd[s] || (d == bd[s] && id < bi[s]);
      if (better) {
        const double td = bd[s]; const int ti = bi[s];
        bd[s] = d; bi[s] = id;
        d = td; id = ti;
      }
    }
  }
  int cnt = 0;
  int* o = nbr + (int64_t)(n0 + c) * k;
#pragma unroll
  for (int s = 0; s < K; ++s) {
    if (s < k) {
      const int v = bi[s] >= 0 ? n0 + bi[s] : -1;
      o[s] = v;
      cnt += v >= 0;
    }
  }
  deg[n0 + c] = cnt;
}

// (nbr, deg, offsets = exclusive scan of deg) -> COO edge_index [2, num_edges] (int64): row 0 neighbour, row 1 centre
__global__ void __launch_bounds__(256) k_knn_to_coo(const int* __restrict__ nbr, const int* __restrict__ deg,
                                                    const int64_t* __restrict__ offsets, int64_t n_total, int k,
                                                    int64_t num_edges, int64_t* __restrict__ edge_index) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_total) return;
  const int d = __ldg(deg + c);
  const int64_t o = __ldg(offsets + c);
  for (int s = 0; s < d; ++s) {
    edge_index[o + s] = __ldg(nbr + c * k + s);
    edge_index[num_edges + o + s] = c;
  }
}

}  // namespace cgc

using namespace cgc;

extern "C" int cgc_radius_graph(const float* pos, const int32_t* start, int32_t num_graphs, int32_t nmax, double r,
                                int32_t k, int loop, int32_t* nbr, int32_t* deg, void* stream) {
  CGC_REQUIRE(pos && start && nbr && deg && num_graphs > 0 && nmax > 0 && r > 0.0, CGC_ERR_ARG,
              "cgc_radius_graph: bad args");
  CGC_REQUIRE(k >= 1 && k <= 32, CGC_ERR_UNSUPPORTED, "cgc_radius_graph: max_num_neighbors must be 1..32 (got %d)", k);
  const size_t smem = (size_t)nmax * sizeof(float2);
  CGC_REQUIRE(smem <= 200 * 1024, CGC_ERR_UNSUPPORTED, "cgc_radius_graph: %d nodes per graph do not fit shared memory",
              nmax);
  dim3 grid((unsigned)ceil_div(nmax, kKnnThreads), (unsigned)num_graphs);
  cudaStream_t st = as_stream(stream);
#define CGC_KNN(KK)                                                                                        \
  do {                                                                                                     \
    auto kern = k_radius_graph<KK>;                                                                        \
    CGC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));          \
    kern<<<grid, kKnnThreads, smem, st>>>(pos, start, r * r, k, loop, nbr, deg);                           \
  } while (0)
  if (k <= 8) CGC_KNN(8);
  else if (k <= 16) CGC_KNN(16);
  else CGC_KNN(32);
#undef CGC_KNN
  CGC_CHECK_LAUNCH("k_radius_graph");
  return 0;
}

extern "C" int cgc_knn_to_coo(const int32_t* nbr, const int32_t* deg, const int64_t* offsets, int64_t n_total,
                              int32_t k, int64_t num_edges, int64_t* edge_index, void* stream) {
  CGC_REQUIRE(nbr && deg && offsets && edge_index && n_total > 0 && k > 0 && num_edges >= 0, CGC_ERR_ARG,
              "cgc_knn_to_coo: bad args");
  k_knn_to_coo<<<(unsigned)ceil_div(n_total, 256), 256, 0, as_stream(stream)>>>(nbr, deg, offsets, n_total, k,
                                                                               num_edges, edge_index);
  CGC_CHECK_LAUNCH("k_knn_to_coo");
  return 0;
}
