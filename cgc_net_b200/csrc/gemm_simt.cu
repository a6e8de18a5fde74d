// fp32 FMA batched GEMM (exact fp32 products): the shape-general contraction of the library.
// C[i] = alpha * op(A[i]) op(B[i]) (+bias) (*row_scale) (*row mask) + beta * C[i]
// Used for every contraction the tcgen05 path does not take (small / unaligned shapes) and as
// the precision = 0 reference of the tensor-core kernel.  Reference call sites: lin of GNN_Module
// (model/network.py:121-124), S^T X / S^T A S of _diff_pool (:206-207), matmul(adj, x) of
// DenseSAGEConv on the dense stages, and their gradients.
#include "common.cuh"
#include "gemm.cuh"

namespace cgc {

constexpr int BK = 16;

// Stage one [BX x BK] operand tile through registers.  KCONTIG: element (x, k) at base[x*ld + k],
// otherwise at base[k*ld + x].  Shared layout: s[k][x] with row stride BX + 4.
template <int BX, int NT, bool KCONTIG>
struct TileLoader {
  static constexpr int kVecs = BX * BK / 4;
  static constexpr int kPerThread = (kVecs + NT - 1) / NT;
  float4 reg[kPerThread];

  __device__ __forceinline__ void load(const float* __restrict__ base, int64_t ld, int x0, int k0, int xdim,
                                       int kend, bool vec_ok) {
#pragma unroll
    for (int u = 0; u < kPerThread; ++u) {
      int f = threadIdx.x + u * NT;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f < kVecs) {
        if (KCONTIG) {
          int x = f / (BK / 4), kq = f - x * (BK / 4);
          int gx = x0 + x, gk = k0 + 4 * kq;
          if (gx < xdim) {
            const float* p = base + (int64_t)gx * ld + gk;
            if (vec_ok && gk + 3 < kend) {
              v = __ldg(reinterpret_cast<const float4*>(p));
            } else {
              if (gk + 0 < kend) v.x = __ldg(p + 0);
              if (gk + 1 < kend) v.y = __ldg(p + 1);
              if (gk + 2 < kend) v.z = __ldg(p + 2);
              if (gk + 3 < kend) v.w = __ldg(p + 3);
            }
          }
        } else {
          int k = f / (BX / 4), xq = f - k * (BX / 4);
          int gk = k0 + k, gx = x0 + 4 * xq;
          if (gk < kend) {
            const float* p = base + (int64_t)gk * ld + gx;
            if (vec_ok && gx + 3 < xdim) {
              v = __ldg(reinterpret_cast<const float4*>(p));
            } else {
              if (gx + 0 < xdim) v.x = __ldg(p + 0);
              if (gx + 1 < xdim) v.y = __ldg(p + 1);
              if (gx + 2 < xdim) v.z = __ldg(p + 2);
              if (gx + 3 < xdim) v.w = __ldg(p + 3);
            }
          }
        }
      }
      reg[u] = v;
    }
  }

  __device__ __forceinline__ void store(float* __restrict__ s) const {
    constexpr int LDS = BX + 4;
#pragma unroll
    for (int u = 0; u < kPerThread; ++u) {
      int f = threadIdx.x + u * NT;
      if (f < kVecs) {
        if (KCONTIG) {
          int x = f / (BK / 4), kq = f - x * (BK / 4);
          s[(4 * kq + 0) * LDS + x] = reg[u].x;
          s[(4 * kq + 1) * LDS + x] = reg[u].y;
          s[(4 * kq + 2) * LDS + x] = reg[u].z;
          s[(4 * kq + 3) * LDS + x] = reg[u].w;
        } else {
          int k = f / (BX / 4), xq = f - k * (BX / 4);
          *reinterpret_cast<float4*>(s + k * LDS + 4 * xq) = reg[u];
        }
      }
    }
  }
};

// thread tile: TMG groups of 4 rows (stride BM/TMG) x TNG groups of 4 columns (stride BN/TNG)
template <int BM, int BN, int TMG, int TNG, bool TA, bool TB>
__global__ void __launch_bounds__((BM / (4 * TMG)) * (BN / (4 * TNG)))
k_gemm_simt(GemmParams p) {
  constexpr int TX = BN / (4 * TNG), TY = BM / (4 * TMG), NT = TX * TY;
  constexpr int LDA = BM + 4, LDB = BN + 4;
  __shared__ __align__(16) float sA[2][BK * LDA];
  __shared__ __align__(16) float sB[2][BK * LDB];

  const int bz = blockIdx.z / p.ksplit, split = blockIdx.z - bz * p.ksplit;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = split * p.kchunk;
  const int kend = min(p.k, kbeg + p.kchunk);
  const float* A = p.a + (int64_t)bz * p.stride_a;
  const float* B = p.b + (int64_t)bz * p.stride_b;
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;

  TileLoader<BM, NT, !TA> la;     // A [m,k]: k-contiguous unless stored transposed
  TileLoader<BN, NT, TB> lb;      // B [k,n]: stored [n,k] (k-contiguous) when trans_b

  float acc[TMG * 4][TNG * 4];
#pragma unroll
  for (int i = 0; i < TMG * 4; ++i)
#pragma unroll
    for (int j = 0; j < TNG * 4; ++j) acc[i][j] = 0.f;

  const int nk = (kend - kbeg + BK - 1) / BK;
  if (nk > 0) {
    la.load(A, p.lda, m0, kbeg, p.m, kend, p.a_vec);
    lb.load(B, p.ldb, n0, kbeg, p.n, kend, p.b_vec);
    la.store(sA[0]);
    lb.store(sB[0]);
  }
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) {
      la.load(A, p.lda, m0, kbeg + (kt + 1) * BK, p.m, kend, p.a_vec);
      lb.load(B, p.ldb, n0, kbeg + (kt + 1) * BK, p.n, kend, p.b_vec);
    }
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TMG * 4], b[TNG * 4];
#pragma unroll
      for (int gi = 0; gi < TMG; ++gi) {
        float4 v = *reinterpret_cast<const float4*>(&sA[buf][kk * LDA + gi * (BM / TMG) + ty * 4]);
        a[gi * 4 + 0] = v.x; a[gi * 4 + 1] = v.y; a[gi * 4 + 2] = v.z; a[gi * 4 + 3] = v.w;
      }
#pragma unroll
      for (int gj = 0; gj < TNG; ++gj) {
        float4 v = *reinterpret_cast<const float4*>(&sB[buf][kk * LDB + gj * (BN / TNG) + tx * 4]);
        b[gj * 4 + 0] = v.x; b[gj * 4 + 1] = v.y; b[gj * 4 + 2] = v.z; b[gj * 4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TMG * 4; ++i)
#pragma unroll
        for (int j = 0; j < TNG * 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      la.store(sA[buf ^ 1]);
      lb.store(sB[buf ^ 1]);
    }
    __syncthreads();
  }

  // epilogue
#pragma unroll
  for (int gi = 0; gi < TMG; ++gi)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = m0 + gi * (BM / TMG) + ty * 4 + i;
      if (row >= p.m) continue;
#pragma unroll
      for (int gj = 0; gj < TNG; ++gj)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = n0 + gj * (BN / TNG) + tx * 4 + j;
          if (col >= p.n) continue;
          float v = p.alpha * acc[gi * 4 + i][gj * 4 + j];
          if (p.ksplit > 1) {
            p.partial[((int64_t)blockIdx.z * p.m + row) * p.n + col] = v;
          } else {
            gemm_store(p, bz, row, col, v);
          }
        }
    }
}

// sum the split-K partials in a fixed order and apply the epilogue
__global__ void __launch_bounds__(256) k_gemm_splitk_reduce(GemmParams p) {
  const int64_t per = (int64_t)p.m * p.n;
  const int64_t total = per * p.batch;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int bz = (int)(idx / per);
    const int64_t rem = idx - (int64_t)bz * per;
    const int row = (int)(rem / p.n), col = (int)(rem - (int64_t)row * p.n);
    float v = 0.f;
    for (int s = 0; s < p.ksplit; ++s) v += p.partial[((int64_t)(bz * p.ksplit + s)) * per + rem];
    gemm_store(p, bz, row, col, v);
  }
}

template <int BM, int BN, int TMG, int TNG>
static int launch_cfg(const GemmParams& p, int ta, int tb, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(p.n, BN), (unsigned)ceil_div(p.m, BM), (unsigned)(p.batch * p.ksplit));
  constexpr int NT = (BM / (4 * TMG)) * (BN / (4 * TNG));
  if (!ta && !tb) k_gemm_simt<BM, BN, TMG, TNG, false, false><<<grid, NT, 0, st>>>(p);
  else if (!ta && tb) k_gemm_simt<BM, BN, TMG, TNG, false, true><<<grid, NT, 0, st>>>(p);
  else if (ta && !tb) k_gemm_simt<BM, BN, TMG, TNG, true, false><<<grid, NT, 0, st>>>(p);
  else k_gemm_simt<BM, BN, TMG, TNG, true, true><<<grid, NT, 0, st>>>(p);
  CGC_CHECK_LAUNCH("k_gemm_simt");
  return 0;
}

int gemm_simt(GemmParams p, int ta, int tb, void* work, int64_t work_bytes, cudaStream_t st) {
  // tile configuration
  int cfg = 0;                       // 0: 128x128, 1: 128x32 (narrow N), 2: 32x128 (narrow M)
  if (p.n <= 64) cfg = 1;
  else if (p.m <= 32) cfg = 2;
  const int bm = cfg == 2 ? 32 : 128, bn = cfg == 1 ? 32 : 128;
  CGC_REQUIRE(ceil_div(p.m, bm) <= 65535, CGC_ERR_UNSUPPORTED, "cgc_gemm: m too large for the grid");
  // split-K when the output grid cannot fill the machine
  const int64_t ctas = ceil_div(p.m, bm) * ceil_div(p.n, bn) * p.batch;
  p.ksplit = 1;
  p.kchunk = p.k;
  if (ctas < 2 * kNumSMs && p.k >= 2048) {
    int64_t ks = ceil_div(4 * kNumSMs, ctas);
    int64_t kmax = p.k / 512;
    if (ks > kmax) ks = kmax;
    if (ks > 256) ks = 256;
    int64_t need = ks * p.batch * (int64_t)p.m * p.n * 4;
    while (ks > 1 && need > work_bytes) { --ks; need = ks * p.batch * (int64_t)p.m * p.n * 4; }
    if (ks > 1) {
      CGC_REQUIRE(work != nullptr, CGC_ERR_WORKSPACE, "cgc_gemm: split-K needs a workspace");
      int64_t chunk = ceil_div(p.k, ks);
      chunk = ceil_div(chunk, BK) * BK;
      p.kchunk = (int)chunk;
      p.ksplit = (int)ceil_div(p.k, chunk);
      p.partial = reinterpret_cast<float*>(work);
    }
  }
  CGC_REQUIRE((int64_t)p.batch * p.ksplit <= 65535, CGC_ERR_UNSUPPORTED, "cgc_gemm: batch*ksplit > 65535");
  int rc;
  if (cfg == 0) rc = launch_cfg<128, 128, 2, 2>(p, ta, tb, st);
  else if (cfg == 1) rc = launch_cfg<128, 32, 2, 1>(p, ta, tb, st);
  else rc = launch_cfg<32, 128, 1, 2>(p, ta, tb, st);
  if (rc) return rc;
  if (p.ksplit > 1) {
    int64_t total = (int64_t)p.m * p.n * p.batch;
    int64_t blocks = ceil_div(total, 256);
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    k_gemm_splitk_reduce<<<(unsigned)blocks, 256, 0, st>>>(p);
    CGC_CHECK_LAUNCH("k_gemm_splitk_reduce");
  }
  return 0;
}

}  // namespace cgc
