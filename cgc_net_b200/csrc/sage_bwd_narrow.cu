// Fused backward of a NARROW SAGE layer epilogue (cout <= 32, cin <= 64): one pass over the rows produces
//   dh  = d loss / d (sa W + b)          (not materialised)
//   dsa = dh W^T                         [rows, cin]
//   dW  = sa^T dh,  db = sum_rows dh     (per-CTA partials, summed in fixed order by a finalize kernel)
// Replaces, for the 20-channel layers, the chain cgc_sage_dh -> cgc_gemm (dW, split-K) -> split-K reduce ->
// cgc_colsum -> finalize -> cgc_gemm (dsa): six launches over the same 9 MB of data.
// Reference arithmetic: backward of PyG 1.2.1 DenseSAGEConv.forward (matmul(out, weight) + bias,
// F.normalize, * mask) followed by nn.ReLU (model/network.py:114-116); formulas in SURVEY.md Appendix B-4.
#include "common.cuh"

namespace cgc {

constexpr int kNbThreads = 256;

// one warp per row; lane c owns output column c (c < cout) and input channels c, c + 32 (< cin)
template <int CINP>
__global__ void __launch_bounds__(kNbThreads)
k_sage_bwd_narrow(const float* __restrict__ gy, int64_t ldgy, const float* __restrict__ sa, int64_t ldsa,
                  const float* __restrict__ w, const float* __restrict__ bias, const float* __restrict__ y,
                  int64_t ldy, const int32_t* __restrict__ num_nodes, int nmax, int64_t rows, int cin, int cout,
                  float* __restrict__ dsa, int64_t lddsa, float* __restrict__ partial) {
  extern __shared__ float smem[];
  float* s_w = smem;                         // [cin][cout]
  float* s_wt = s_w + cin * cout;            // [cout][CINP + 1]  (transposed, padded: conflict-free by lane = k)
  float* s_b = s_wt + cout * (CINP + 1);     // [cout]
  float* s_red = s_b + 32;                   // [8 warps][CINP * 32 + 32]  block reduction of dW / db
  for (int i = threadIdx.x; i < cin * cout; i += kNbThreads) {
    const float v = w[i];
    s_w[i] = v;
    const int k = i / cout, c = i - k * cout;
    s_wt[c * (CINP + 1) + k] = v;
  }
  for (int i = threadIdx.x; i < 32; i += kNbThreads) s_b[i] = (bias && i < cout) ? bias[i] : 0.f;
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t gwarp = (int64_t)blockIdx.x * (kNbThreads / 32) + warp;
  const int64_t nwarps = (int64_t)gridDim.x * (kNbThreads / 32);
  const int c = lane < cout ? lane : 0;

  float dw_acc[CINP];                        // dW[k][lane] partial over this warp's rows
#pragma unroll
  for (int k = 0; k < CINP; ++k) dw_acc[k] = 0.f;
  float db_acc = 0.f;

  for (int64_t r = gwarp; r < rows; r += nwarps) {
    const float* srow = sa + r * ldsa;
    const float a0 = lane < cin ? __ldg(srow + lane) : 0.f;
    const float a1 = (CINP > 32 && 32 + lane < cin) ? __ldg(srow + 32 + lane) : 0.f;
    const bool valid = row_valid(num_nodes, nmax, r);
    float g = 0.f;
    if (lane < cout && valid) {
      const float yv = __ldg(y + r * ldy + lane);
      const float gv = __ldg(gy + r * ldgy + lane);
      g = yv > 0.f ? gv : 0.f;               // ReLU by output sign, row mask
    }
    // h = sa W + b (recomputed), hn = h / max(|h|, eps)
    float h = lane < cout ? s_b[lane] : 0.f;
#pragma unroll
    for (int k = 0; k < CINP; ++k) {
      if (k < cin) {
        const float a = (k < 32) ? __shfl_sync(0xffffffffu, a0, k) : __shfl_sync(0xffffffffu, a1, k - 32);
        h = fmaf(a, s_w[k * cout + c], h);
      }
    }
    if (lane >= cout) h = 0.f;
    const float nrm = sqrtf(warp_sum(h * h));
    const float inv = 1.0f / fmaxf(nrm, kL2NormEps);
    const float hn = h * inv;
    const float dot = warp_sum(g * hn);
    float d = (nrm >= kL2NormEps) ? (g - hn * dot) * inv : g * inv;
    if (lane >= cout) d = 0.f;
    db_acc += d;
    // dW[k][lane] += sa[r][k] * dh[r][lane]
#pragma unroll
    for (int k = 0; k < CINP; ++k) {
      if (k < cin) {
        const float a = (k < 32) ? __shfl_sync(0xffffffffu, a0, k) : __shfl_sync(0xffffffffu, a1, k - 32);
        dw_acc[k] = fmaf(a, d, dw_acc[k]);
      }
    }
    // dsa[r][k] = sum_c dh[r][c] W[k][c]   (lane = k, and k + 32)
    float o0 = 0.f, o1 = 0.f;
    for (int cc = 0; cc < cout; ++cc) {
      const float dc = __shfl_sync(0xffffffffu, d, cc);
      o0 = fmaf(dc, s_wt[cc * (CINP + 1) + lane], o0);
      if (CINP > 32) o1 = fmaf(dc, s_wt[cc * (CINP + 1) + 32 + lane], o1);
    }
    if (lane < cin) dsa[r * lddsa + lane] = o0;
    if (CINP > 32 && 32 + lane < cin) dsa[r * lddsa + 32 + lane] = o1;
  }

  // block reduction (fixed order) -> one partial per CTA: [cin * 32 + 32]
  float* mine = s_red + warp * (CINP * 32 + 32);
#pragma unroll
  for (int k = 0; k < CINP; ++k) mine[k * 32 + lane] = dw_acc[k];
  mine[CINP * 32 + lane] = db_acc;
  __syncthreads();
  float* out = partial + (int64_t)blockIdx.x * (CINP * 32 + 32);
  for (int i = threadIdx.x; i < CINP * 32 + 32; i += kNbThreads) {
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < kNbThreads / 32; ++wv) s += s_red[wv * (CINP * 32 + 32) + i];
    out[i] = s;
  }
}

// d_w[k][c] = sum_blocks partial[b][k*32 + c];  d_b[c] = sum_blocks partial[b][CINP*32 + c]
__global__ void __launch_bounds__(256)
k_sage_bwd_narrow_finalize(const float* __restrict__ partial, int nblocks, int cinp, int cin, int cout,
                           float* __restrict__ d_w, float* __restrict__ d_b) {
  const int lane = threadIdx.x & 31;
  const int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;          // one warp per output element
  const int n_items = cin * cout + cout;
  if (item >= n_items) return;
  int src;
  if (item < cin * cout) {
    const int k = item / cout, c = item - k * cout;
    src = k * 32 + c;
  } else {
    src = cinp * 32 + (item - cin * cout);
  }
  const int stride = cinp * 32 + 32;
  double s = 0.0;
  for (int b = lane; b < nblocks; b += 32) s += (double)partial[(int64_t)b * stride + src];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    if (item < cin * cout) d_w[item] = (float)s;
    else if (d_b) d_b[item - cin * cout] = (float)s;
  }
}

}  // namespace cgc

using namespace cgc;

extern "C" int cgc_sage_layer_bwd_narrow(const float* gy, int64_t ldgy, const float* sa, int64_t ldsa,
                                         const float* w, const float* b, const float* y, int64_t ldy,
                                         const int32_t* num_nodes, int32_t nmax, int64_t rows, int32_t cin,
                                         int32_t cout, float* dsa, int64_t lddsa, float* d_w, float* d_b,
                                         void* work, int64_t work_bytes, void* stream) {
  CGC_REQUIRE(gy && sa && w && y && dsa && d_w && work && rows > 0 && cin > 0 && cout > 0, CGC_ERR_ARG,
              "cgc_sage_layer_bwd_narrow: bad args");
  CGC_REQUIRE(cout <= 32 && cin <= 64, CGC_ERR_UNSUPPORTED,
              "cgc_sage_layer_bwd_narrow: needs cout <= 32 and cin <= 64 (got %d, %d)", cout, cin);
  CGC_REQUIRE(num_nodes == nullptr || nmax > 0, CGC_ERR_ARG, "cgc_sage_layer_bwd_narrow: mask without nmax");
  const int cinp = cin <= 16 ? 16 : (cin <= 32 ? 32 : 64);
  int64_t blocks = ceil_div(rows, kNbThreads / 32);
  if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
  const int64_t stride = cinp * 32 + 32;
  CGC_REQUIRE(blocks * stride * 4 <= work_bytes, CGC_ERR_WORKSPACE, "cgc_sage_layer_bwd_narrow: workspace too small");
  const size_t smem = sizeof(float) * (size_t)(cin * cout + cout * (cinp + 1) + 32 + (kNbThreads / 32) * stride);
  cudaStream_t st = as_stream(stream);
  float* partial = reinterpret_cast<float*>(work);
#define CGC_NB_CASE(N)                                                                                          \
  case N: {                                                                                                     \
    auto kern = k_sage_bwd_narrow<N>;                                                                           \
    CGC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));               \
    kern<<<(unsigned)blocks, kNbThreads, smem, st>>>(gy, ldgy, sa, ldsa, w, b, y, ldy, num_nodes, nmax, rows, cin, \
                                                     cout, dsa, lddsa, partial);                                \
  } break;
  switch (cinp) {
    CGC_NB_CASE(16)
    CGC_NB_CASE(32)
    CGC_NB_CASE(64)
  }
#undef CGC_NB_CASE
  CGC_CHECK_LAUNCH("k_sage_bwd_narrow");
  const int n_items = cin * cout + cout;
  k_sage_bwd_narrow_finalize<<<(unsigned)ceil_div(n_items, 8), 256, 0, st>>>(partial, (int)blocks, cinp, cin, cout,
                                                                           d_w, d_b);
  CGC_CHECK_LAUNCH("k_sage_bwd_narrow_finalize");
  return 0;
}
