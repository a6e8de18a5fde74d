// Cell-graph construction on the GPU: radius_graph(pos, r, batch, loop=True, max_num_neighbors=k) of
// torch_cluster 1.4.2 as the reference calls it (dataflow/data.py:246,255,297,348): for every centre node the
// <= k nearest nodes of the same graph within distance r (the centre itself first, at distance 0), returned as
// COO pairs (edge_index[0] = neighbour, edge_index[1] = centre).  The CPU implementation behind the reference is
// scipy's cKDTree.query(k, distance_upper_bound=r); here one thread owns one centre, scans its graph's positions
// from shared memory and keeps the k best (distance, index) pairs in registers.  Distances are evaluated in
// double precision on the float positions (as cKDTree does), ties broken by the lower node index.
#include <limits.h>

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

// ------------------------------------------------------------------ farthest-point node sampling
// FarthestSampler.__call__ (reference common/utils.py:187-197) on the distance matrix that
// construct_feature_graph.py:17-24 stores: d(i, j) = (int16) sqrt((xi-xj)^2 + (yi-yj)^2) (truncation).  The
// reference walks a [N, N] int16 matrix loaded from disk; here one CTA owns one graph, keeps the running
// min-distance vector in shared memory and evaluates the distances from the positions on the fly:
//   sel_0 = first[g];  mind = d(sel_0, .);  sel_i = argmax(mind) (first maximum, like np.argmax);
//   mind = min(mind, d(sel_i, .)).
namespace cgc {

constexpr int kFpsThreads = 256;

__device__ __forceinline__ int fps_dist(float2 a, float2 b) {
  const double dx = (double)a.x - (double)b.x, dy = (double)a.y - (double)b.y;
  return (int)(short)(long long)sqrt(dx * dx + dy * dy);       // astype(np.int16) of the float64 distance
}

__global__ void __launch_bounds__(kFpsThreads)
k_farthest_point_sample(const float* __restrict__ pos, const int* __restrict__ start, const int* __restrict__ first,
                        const int* __restrict__ out_start, int* __restrict__ out) {
  extern __shared__ unsigned char fps_smem[];
  const int g = blockIdx.x;
  const int n0 = __ldg(start + g), n = __ldg(start + g + 1) - n0;
  const int o0 = __ldg(out_start + g), k = __ldg(out_start + g + 1) - o0;
  if (n <= 0 || k <= 0) return;
  float2* s_pos = reinterpret_cast<float2*>(fps_smem);
  int* s_min = reinterpret_cast<int*>(s_pos + n);
  __shared__ int s_val[kFpsThreads / 32], s_idx[kFpsThreads / 32], s_sel;
  const int tid = threadIdx.x;
  for (int i = tid; i < n; i += kFpsThreads)
    s_pos[i] = make_float2(__ldg(pos + 2 * (int64_t)(n0 + i)), __ldg(pos + 2 * (int64_t)(n0 + i) + 1));
  int sel = min(max(__ldg(first + g), 0), n - 1);
  __syncthreads();
  if (tid == 0) out[o0] = sel;
  for (int i = tid; i < n; i += kFpsThreads) s_min[i] = fps_dist(s_pos[sel], s_pos[i]);
  for (int it = 1; it < k; ++it) {
    __syncthreads();
    // arg max with the lowest index on ties
    int bv = INT_MIN, bi = 0x7fffffff;
    for (int i = tid; i < n; i += kFpsThreads) {
      const int v = s_min[i];
      if (v > bv) { bv = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const int ov = __shfl_xor_sync(0xffffffffu, bv, o), oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((tid & 31) == 0) { s_val[tid >> 5] = bv; s_idx[tid >> 5] = bi; }
    __syncthreads();
    if (tid == 0) {
      int v = s_val[0], ix = s_idx[0];
      for (int w = 1; w < kFpsThreads / 32; ++w)
        if (s_val[w] > v || (s_val[w] == v && s_idx[w] < ix)) { v = s_val[w]; ix = s_idx[w]; }
      s_sel = ix;
      out[o0 + it] = ix;
    }
    __syncthreads();
    sel = s_sel;
    const float2 ps = s_pos[sel];
    for (int i = tid; i < n; i += kFpsThreads) s_min[i] = min(s_min[i], fps_dist(ps, s_pos[i]));
  }
}

}  // namespace cgc

extern "C" int cgc_farthest_point_sample(const float* pos, const int32_t* start, const int32_t* first,
                                         const int32_t* out_start, int32_t num_graphs, int32_t nmax, int32_t* out,
                                         void* stream) {
  CGC_REQUIRE(pos && start && first && out_start && out && num_graphs > 0 && nmax > 0, CGC_ERR_ARG,
              "cgc_farthest_point_sample: bad args");
  const size_t smem = (size_t)nmax * (sizeof(float2) + sizeof(int));
  CGC_REQUIRE(smem <= 200 * 1024, CGC_ERR_UNSUPPORTED,
              "cgc_farthest_point_sample: %d nodes per graph do not fit shared memory", nmax);
  CGC_CUDA(cudaFuncSetAttribute(cgc::k_farthest_point_sample, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cgc::k_farthest_point_sample<<<(unsigned)num_graphs, cgc::kFpsThreads, smem, as_stream(stream)>>>(pos, start, first,
                                                                                                   out_start, out);
  CGC_CHECK_LAUNCH("k_farthest_point_sample");
  return 0;
}
