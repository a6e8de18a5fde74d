// Fused backward of a NARROW SAGE layer epilogue (cout <= 32, cin <= 64): one pass over the rows produces
//   dh  = d loss / d (sa W + b)          (not materialised)
//   dsa = dh W^T                         [rows, cin]
//   dW  = sa^T dh,  db = sum_rows dh     (per-CTA partials, summed in fixed order by a finalize kernel)
// Replaces, for the 20-channel layers, the chain cgc_sage_dh -> cgc_gemm (dW, split-K) -> split-K reduce ->
// cgc_colsum -> finalize -> cgc_gemm (dsa): six launches over the same 9 MB of data.
//
// Work decomposition: ONE THREAD PER ROW for the row-local math (h = sa W + b recomputed, the L2-normalise /
// ReLU / mask backward, dsa = dh W^T) with W read from shared memory as warp-wide broadcasts, so a row costs
// ~2 cin cout FMAs and cin cout / 2 broadcast LDS.128 per THREAD (no shuffles).  dW / db, the only cross-row
// quantity, is a small register-tiled contraction over the CTA's 256-row tile: the rows' sa (extended by a
// column of ones, which makes db the last row of dW) and dh are staged in shared memory and each thread
// owns one 4x4 block of dW for a subset of the rows.
//
// Reference arithmetic: backward of PyG 1.2.1 DenseSAGEConv.forward (matmul(out, weight) + bias,
// F.normalize, * mask) followed by nn.ReLU (model/network.py:114-116); formulas in SURVEY.md Appendix B-4.
#include "common.cuh"

namespace cgc {

constexpr int kNbThreads = 256;              // = rows per tile
constexpr int kNbWStride = 32;               // W staged as [k][32], zero padded

// smallest multiple of 4 that is >= n and whose quarter is odd: float4 accesses by 8 consecutive threads
// at this row stride hit 8 distinct 16-byte bank groups
__host__ __device__ inline int nb_stride(int n) {
  int s = (n + 3) & ~3;
  return ((s >> 2) & 1) ? s : s + 4;
}

// C4 = number of float4 column groups held in registers (cout <= 4 * C4)
template <int C4>
__global__ void __launch_bounds__(kNbThreads, 3)
k_sage_bwd_narrow(const float* __restrict__ gy, int64_t ldgy, const float* __restrict__ sa, int64_t ldsa,
                  const float* __restrict__ w, const float* __restrict__ bias, const float* __restrict__ y,
                  int64_t ldy, const int32_t* __restrict__ num_nodes, int nmax, int64_t rows, int cin, int cout,
                  float* __restrict__ dsa, int64_t lddsa, float* __restrict__ partial, int cinp, int vec, int act) {
  extern __shared__ __align__(16) float smem[];
  const ActGrad agrad(act);
  const int cin_r = (cin + 3) & ~3;          // W rows staged (zero padded)
  const int cin_e = (cin + 4) & ~3;          // sa columns of the dW contraction: cin, the ones column, zero pad
  const int SA = nb_stride(cin + 1), SD = nb_stride(cout);
  float* s_w = smem;                         // [cin_r][32]
  float* s_b = s_w + cin_r * kNbWStride;     // [32]
  float* s_sa = s_b + 32;                    // [256][SA]
  float* s_d = s_sa + kNbThreads * SA;       // [256][SD]
  float* s_red = s_sa;                       // aliased after the last tile: [groups][ntiles][16]
  const int t = threadIdx.x;
  for (int i = t; i < cin_r * kNbWStride; i += kNbThreads) {
    const int k = i >> 5, c = i & 31;
    s_w[i] = (k < cin && c < cout) ? w[k * cout + c] : 0.f;
  }
  if (t < 32) s_b[t] = (bias && t < cout) ? bias[t] : 0.f;

  // dW tiling: 4x4 blocks, (cin_e/4) x ceil(cout/4) of them, replicated over `groups` row subsets
  const int tcn = (cout + 3) >> 2;
  const int ntiles = (cin_e >> 2) * tcn;
  const int groups = kNbThreads / ntiles;
  const int grp = t / ntiles, tile = t - grp * ntiles;
  const int tk = tile / tcn, tc = tile - tk * tcn;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float* my_sa = s_sa + t * SA;
  float* my_d = s_d + t * SD;
  const int64_t ntile_rows = (rows + kNbThreads - 1) / kNbThreads;
  for (int64_t rt = blockIdx.x; rt < ntile_rows; rt += gridDim.x) {
    const int64_t r = rt * kNbThreads + t;
    const bool in_range = r < rows;
    __syncthreads();                          // W staged / previous tile's contraction done
    // ---- stage this row of sa (+ ones column)
    if (in_range) {
      const float* srow = sa + r * ldsa;
      if (vec) {
        for (int k4 = 0; k4 < (cin >> 2); ++k4)
          *reinterpret_cast<float4*>(my_sa + 4 * k4) = __ldg(reinterpret_cast<const float4*>(srow) + k4);
      } else {
        for (int k = 0; k < cin; ++k) my_sa[k] = __ldg(srow + k);
      }
      my_sa[cin] = 1.f;
      for (int k = cin + 1; k < cin_e; ++k) my_sa[k] = 0.f;
    } else {
      for (int k = 0; k < cin_e; ++k) my_sa[k] = 0.f;
    }
    // ---- h = sa W + b
    float h[4 * C4];
#pragma unroll
    for (int c = 0; c < 4 * C4; ++c) h[c] = s_b[c];
    for (int k4 = 0; k4 < (cin_r >> 2); ++k4) {
      const float4 a4 = *reinterpret_cast<const float4*>(my_sa + 4 * k4);   // column cin (= 1) meets a zero W row
      const float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float4* wr = reinterpret_cast<const float4*>(s_w + (4 * k4 + kk) * kNbWStride);
#pragma unroll
        for (int c4 = 0; c4 < C4; ++c4) {
          const float4 wv = wr[c4];
          h[4 * c4 + 0] = fmaf(av[kk], wv.x, h[4 * c4 + 0]);
          h[4 * c4 + 1] = fmaf(av[kk], wv.y, h[4 * c4 + 1]);
          h[4 * c4 + 2] = fmaf(av[kk], wv.z, h[4 * c4 + 2]);
          h[4 * c4 + 3] = fmaf(av[kk], wv.w, h[4 * c4 + 3]);
        }
      }
    }
    float n2 = 0.f;
#pragma unroll
    for (int c = 0; c < 4 * C4; ++c) n2 = fmaf(h[c], h[c], n2);            // padded columns are exactly 0
    const float nrm = sqrtf(n2);
    const float inv = 1.0f / fmaxf(nrm, kL2NormEps);
    // ---- g = gy * [y > 0] * mask (parked in this row's s_d slot);  dh = (g - hn (g . hn)) / |h|
    const bool valid = in_range && row_valid(num_nodes, nmax, r);
    const float* yrow = y + r * ldy;
    const float* grow = gy + r * ldgy;
    float dot = 0.f;
#pragma unroll
    for (int c4 = 0; c4 < C4; ++c4) {
      if (4 * c4 < SD) {
        float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid && 4 * c4 < cout) {
          if (vec) {
            const float4 yv = __ldg(reinterpret_cast<const float4*>(yrow) + c4);
            const float4 gv = __ldg(reinterpret_cast<const float4*>(grow) + c4);
            g4.x = gv.x * agrad(yv.x);
            g4.y = gv.y * agrad(yv.y);
            g4.z = gv.z * agrad(yv.z);
            g4.w = gv.w * agrad(yv.w);
          } else {
            float gs[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int c = 4 * c4 + j;
              gs[j] = c < cout ? __ldg(grow + c) * agrad(__ldg(yrow + c)) : 0.f;
            }
            g4 = make_float4(gs[0], gs[1], gs[2], gs[3]);
          }
        }
        *reinterpret_cast<float4*>(my_d + 4 * c4) = g4;
        h[4 * c4 + 0] *= inv;                 // hn
        h[4 * c4 + 1] *= inv;
        h[4 * c4 + 2] *= inv;
        h[4 * c4 + 3] *= inv;
        dot = fmaf(g4.x, h[4 * c4 + 0], dot);
        dot = fmaf(g4.y, h[4 * c4 + 1], dot);
        dot = fmaf(g4.z, h[4 * c4 + 2], dot);
        dot = fmaf(g4.w, h[4 * c4 + 3], dot);
      }
    }
    const bool regular = nrm >= kL2NormEps;
#pragma unroll
    for (int c4 = 0; c4 < C4; ++c4) {
      if (4 * c4 < SD) {
        const float4 g4 = *reinterpret_cast<const float4*>(my_d + 4 * c4);
        h[4 * c4 + 0] = regular ? (g4.x - h[4 * c4 + 0] * dot) * inv : g4.x * inv;   // h := dh
        h[4 * c4 + 1] = regular ? (g4.y - h[4 * c4 + 1] * dot) * inv : g4.y * inv;
        h[4 * c4 + 2] = regular ? (g4.z - h[4 * c4 + 2] * dot) * inv : g4.z * inv;
        h[4 * c4 + 3] = regular ? (g4.w - h[4 * c4 + 3] * dot) * inv : g4.w * inv;
        *reinterpret_cast<float4*>(my_d + 4 * c4) = make_float4(h[4 * c4], h[4 * c4 + 1], h[4 * c4 + 2], h[4 * c4 + 3]);
      }
    }
    // ---- dsa[r][k] = sum_c dh[c] W[k][c]
    if (in_range) {
      float* orow = dsa + r * lddsa;
      for (int k4 = 0; k4 < (cin_r >> 2); ++k4) {
        float o[4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const float4* wr = reinterpret_cast<const float4*>(s_w + (4 * k4 + kk) * kNbWStride);
          float s = 0.f;
#pragma unroll
          for (int c4 = 0; c4 < C4; ++c4) {
            const float4 wv = wr[c4];
            s = fmaf(h[4 * c4 + 0], wv.x, s);
            s = fmaf(h[4 * c4 + 1], wv.y, s);
            s = fmaf(h[4 * c4 + 2], wv.z, s);
            s = fmaf(h[4 * c4 + 3], wv.w, s);
          }
          o[kk] = s;
        }
        if (vec) {
          *reinterpret_cast<float4*>(orow + 4 * k4) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            if (4 * k4 + kk < cin) orow[4 * k4 + kk] = o[kk];
        }
      }
    }
    __syncthreads();
    // ---- dW (and db as row cin): 4x4 register blocks over this tile's rows
    if (grp < groups) {
      const float* pa = s_sa + 4 * tk;
      const float* pd = s_d + 4 * tc;
#pragma unroll 4
      for (int rr = grp; rr < kNbThreads; rr += groups) {
        const float4 a4 = *reinterpret_cast<const float4*>(pa + rr * SA);
        const float4 d4 = *reinterpret_cast<const float4*>(pd + rr * SD);
        const float av[4] = {a4.x, a4.y, a4.z, a4.w};
        const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], dv[j], acc[i][j]);
      }
    }
  }

  // ---- fixed-order reduction over the row groups -> one partial per CTA: [cinp * 32 + 32]
  __syncthreads();
  if (grp < groups) {
    float* mine = s_red + (grp * ntiles + tile) * 16;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<float4*>(mine + 4 * i) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
  __syncthreads();
  float* out = partial + (int64_t)blockIdx.x * (cinp * 32 + 32);
  for (int idx = t; idx < ntiles * 16; idx += kNbThreads) {
    const int tl = idx >> 4, e = idx & 15;
    float s = 0.f;
    for (int gq = 0; gq < groups; ++gq) s += s_red[(gq * ntiles + tl) * 16 + e];
    const int k = 4 * (tl / tcn) + (e >> 2), c = 4 * (tl % tcn) + (e & 3);
    if (c < cout) {
      if (k < cin) out[k * 32 + c] = s;
      else if (k == cin) out[cinp * 32 + c] = s;
    }
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
                                         int32_t cout, int act, float* dsa, int64_t lddsa, float* d_w, float* d_b,
                                         void* work, int64_t work_bytes, void* stream) {
  CGC_REQUIRE(gy && sa && w && y && dsa && d_w && work && rows > 0 && cin > 0 && cout > 0, CGC_ERR_ARG,
              "cgc_sage_layer_bwd_narrow: bad args");
  CGC_REQUIRE(cout <= 32 && cin <= 64, CGC_ERR_UNSUPPORTED,
              "cgc_sage_layer_bwd_narrow: needs cout <= 32 and cin <= 64 (got %d, %d)", cout, cin);
  CGC_REQUIRE(num_nodes == nullptr || nmax > 0, CGC_ERR_ARG, "cgc_sage_layer_bwd_narrow: mask without nmax");
  CGC_REQUIRE(act >= CGC_ACT_RELU && act <= CGC_ACT_LEAKY_RELU, CGC_ERR_ARG,
              "cgc_sage_layer_bwd_narrow: unknown activation %d", act);
  const int cinp = cin <= 16 ? 16 : (cin <= 32 ? 32 : 64);
  const int c4 = cout <= 8 ? 2 : (cout <= 20 ? 5 : 8);
  const int cin_r = (cin + 3) & ~3;
  const int sa_stride = nb_stride(cin + 1), sd_stride = nb_stride(cout);
  size_t tile_floats = (size_t)kNbThreads * (sa_stride + sd_stride);
  if (tile_floats < (size_t)kNbThreads * 16) tile_floats = (size_t)kNbThreads * 16;      // reduction scratch
  const size_t smem = sizeof(float) * ((size_t)cin_r * kNbWStride + 32 + tile_floats);
  const int per_sm = smem + 1024 <= 75 * 1024 ? 3 : (smem + 1024 <= 113 * 1024 ? 2 : 1);
  int64_t blocks = ceil_div(rows, kNbThreads);
  if (blocks > (int64_t)kNumSMs * per_sm) blocks = (int64_t)kNumSMs * per_sm;
  const int64_t stride = cinp * 32 + 32;
  CGC_REQUIRE(blocks * stride * 4 <= work_bytes, CGC_ERR_WORKSPACE, "cgc_sage_layer_bwd_narrow: workspace too small");
  const int vec = (aligned16(gy) && aligned16(sa) && aligned16(y) && aligned16(dsa) && ldgy % 4 == 0 &&
                   ldsa % 4 == 0 && ldy % 4 == 0 && lddsa % 4 == 0 && cin % 4 == 0 && cout % 4 == 0)
                      ? 1 : 0;
  cudaStream_t st = as_stream(stream);
  float* partial = reinterpret_cast<float*>(work);
#define CGC_NB_CASE(N)                                                                                          \
  case N: {                                                                                                     \
    auto kern = k_sage_bwd_narrow<N>;                                                                           \
    CGC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));               \
    kern<<<(unsigned)blocks, kNbThreads, smem, st>>>(gy, ldgy, sa, ldsa, w, b, y, ldy, num_nodes, nmax, rows, cin, \
                                                     cout, dsa, lddsa, partial, cinp, vec, act);                \
  } break;
  switch (c4) {
    CGC_NB_CASE(2)
    CGC_NB_CASE(5)
    CGC_NB_CASE(8)
  }
#undef CGC_NB_CASE
  CGC_CHECK_LAUNCH("k_sage_bwd_narrow");
  const int n_items = cin * cout + cout;
  k_sage_bwd_narrow_finalize<<<(unsigned)ceil_div(n_items, 8), 256, 0, st>>>(partial, (int)blocks, cinp, cin, cout,
                                                                           d_w, d_b);
  CGC_CHECK_LAUNCH("k_sage_bwd_narrow_finalize");
  return 0;
}
