// Row-wise operators of the SAGE layer and the diff-pool:
//   cgc_sage_epilogue_fwd / cgc_sage_dh   (sa @ W + b, L2-normalise, mask, ReLU and its gradient)
//   cgc_softmax_fwd / cgc_softmax_bwd     (cluster assignment, network.py:200-205)
//   cgc_scale_rows_bwd, cgc_row_inv_clamped_sum, cgc_add_rs_grad, cgc_dense_renorm_fwd/bwd
// Reference: PyG 1.2.1 DenseSAGEConv.forward (called at model/network.py:114-116), nn.ReLU(inplace)
// (network.py:87), _diff_pool (:194-208), _re_norm_adj (:183-191).
#include <string.h>

#include "common.cuh"

namespace cgc {

// =================================================================== narrow epilogue (cout <= 32)
// one warp per row; lane c owns output column c.  W, b live in shared memory.
template <bool BWD>
__global__ void __launch_bounds__(256)
k_epi_narrow(const float* __restrict__ sa, int64_t ldsa, const float* __restrict__ w, const float* __restrict__ bias,
             const int32_t* __restrict__ num_nodes, int nmax, int64_t rows, int cin, int cout,
             float* __restrict__ y, int64_t ldy, const float* __restrict__ gy, int64_t ldgy,
             float* __restrict__ dh, int64_t lddh, int act) {
  extern __shared__ float s_w[];          // [cin][cout] then bias[cout]
  const ActGrad agrad(act);
  float* s_b = s_w + cin * cout;
  for (int i = threadIdx.x; i < cin * cout; i += blockDim.x) s_w[i] = w[i];
  for (int i = threadIdx.x; i < cout; i += blockDim.x) s_b[i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < rows; r += nwarps) {
    const float* srow = sa + r * ldsa;
    float a0 = lane < cin ? __ldg(srow + lane) : 0.f;
    float a1 = (32 + lane) < cin ? __ldg(srow + 32 + lane) : 0.f;
    float h = (lane < cout) ? s_b[lane] : 0.f;
    const int c = lane < cout ? lane : 0;
    for (int k = 0; k < cin; ++k) {
      float a = (k < 32) ? __shfl_sync(0xffffffffu, a0, k) : __shfl_sync(0xffffffffu, a1, k - 32);
      h = fmaf(a, s_w[k * cout + c], h);
    }
    if (lane >= cout) h = 0.f;
    const float nrm = sqrtf(warp_sum(h * h));
    const float inv = 1.0f / fmaxf(nrm, kL2NormEps);
    const bool valid = row_valid(num_nodes, nmax, r);
    if (!BWD) {
      if (lane < cout) y[r * ldy + lane] = valid ? act_fwd(act, h * inv) : 0.f;
    } else {
      float hn = h * inv;
      float g = 0.f;
      if (lane < cout && valid) g = gy[r * ldgy + lane] * agrad(__ldg(y + r * ldy + lane));
      const float dot = warp_sum(g * hn);
      float d = (nrm >= kL2NormEps) ? (g - hn * dot) * inv : g * inv;
      if (lane < cout) dh[r * lddh + lane] = d;
    }
  }
}

// =================================================================== wide epilogue (cout > 32)
// CTA of 256 threads processes RT rows at a time; thread t owns columns t + 256*j (j < NC).
constexpr int kEpiThreads = 256;
constexpr int kEpiRows = 8;

// Backward with the BatchNorm backward folded in (bn.mean != null): `gy` is then the gradient w.r.t. the BatchNorm
// OUTPUT and the kernel forms, per element, gy_bn = P gx - R y - Q with the per-column coefficients
//   P = gamma invstd,  R = gamma invstd^2 d_gamma / rows,  Q = gamma invstd d_beta / rows - R mean     (training)
//   P = gamma invstd,  R = Q = 0                                                                       (eval)
// staged in shared memory once per CTA -- the separate pass that read gx and y and wrote gy (0.29 ms, 1.57 GB at the
// BASELINE shape) disappears.
struct EpiBn {
  const float* mean;
  const float* invstd;
  const float* gamma;
  const float* d_gamma;
  const float* d_beta;
  float inv_rows;
  int training;
};

template <int NC, bool BWD>
__global__ void __launch_bounds__(kEpiThreads, NC == 1 ? 4 : (NC == 2 ? 3 : 2))
k_epi_wide(const float* __restrict__ sa, int64_t ldsa, const float* __restrict__ w, const float* __restrict__ bias,
           const int32_t* __restrict__ num_nodes, int nmax, int64_t rows, int cin, int cout,
           float* __restrict__ y, int64_t ldy, const float* __restrict__ gy, int64_t ldgy,
           float* __restrict__ dh, int64_t lddh, int act, double* __restrict__ partial, const EpiBn bn) {
  extern __shared__ float smem[];
  const ActGrad agrad(act);
  // forward with BatchNorm statistics (partial != null): every thread owns its columns in the CTA, so the column
  // sums of y and y^2 over the CTA's tiles accumulate in registers and leave as one [2][cout] partial per CTA
  float st1[NC], st2[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) { st1[j] = 0.f; st2[j] = 0.f; }
  float* s_w = smem;                                 // [cin][cout]
  float* s_a = s_w + cin * cout;                     // [cin][kEpiRows]  (transposed row tile)
  float* s_red = s_a + cin * kEpiRows;               // [8 warps][kEpiRows]
  int* s_valid = reinterpret_cast<int*>(s_red + (kEpiThreads / 32) * kEpiRows);   // [kEpiRows] row-mask flags of the tile
  float* s_bn = reinterpret_cast<float*>(s_valid + kEpiRows);                     // [3][cout] P, R, Q (BWD with bn)
  const bool with_bn = BWD && bn.mean != nullptr;
  for (int i = threadIdx.x; i < cin * cout; i += kEpiThreads) s_w[i] = w[i];
  if (with_bn) {
    for (int c = threadIdx.x; c < cout; c += kEpiThreads) {
      const float is = bn.invstd[c], ga = bn.gamma[c];
      const float pc = ga * is;
      const float rc = bn.training ? pc * is * bn.d_gamma[c] * bn.inv_rows : 0.f;
      s_bn[c] = pc;
      s_bn[cout + c] = rc;
      s_bn[2 * cout + c] = bn.training ? pc * bn.d_beta[c] * bn.inv_rows - rc * bn.mean[c] : 0.f;
    }
  }
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  float bj[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    int c = t + kEpiThreads * j;
    bj[j] = (bias && c < cout) ? bias[c] : 0.f;
  }
  const int64_t ntiles = (rows + kEpiRows - 1) / kEpiRows;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t r0 = tile * kEpiRows;
    if constexpr (BWD) {
      // The gradient and activation rows of this CTA's NEXT tile are pulled into L2 now, one 128-byte line per
      // thread and matrix: when the tile is reached its 80 loads per thread hit L2 instead of stalling a whole
      // DRAM round trip behind the FMA phase (no registers to spare for a software pipeline here).
      const int64_t rn = (tile + gridDim.x) * kEpiRows;
      const int lines = (cout + 31) >> 5;
      for (int i = t; i < kEpiRows * lines; i += kEpiThreads) {
        const int rr = i / lines, ch = i - rr * lines;
        if (rn + rr < rows) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(y + (rn + rr) * ldy + ch * 32));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(gy + (rn + rr) * ldgy + ch * 32));
        }
      }
    }
    __syncthreads();                                 // previous tile done with s_a / s_red (and s_w loaded)
    for (int i = t; i < cin * kEpiRows; i += kEpiThreads) {
      int rr = i / cin, k = i - rr * cin;
      s_a[k * kEpiRows + rr] = (r0 + rr < rows) ? __ldg(sa + (r0 + rr) * ldsa + k) : 0.f;
    }
    // the row mask once per tile (a 64-bit division per row: 8 threads do it instead of every thread 8 times)
    if (t < kEpiRows) s_valid[t] = (r0 + t < rows && row_valid(num_nodes, nmax, r0 + t)) ? 1 : 0;
    __syncthreads();
    uint32_t vmask = 0;                              // flags in a register: the row loops below stay branch-free
#pragma unroll
    for (int rr = 0; rr < kEpiRows; ++rr) vmask |= (uint32_t)s_valid[rr] << rr;
    // h = a W + b in packed f32x2 arithmetic: one register pair holds the accumulators of two ROWS of the tile for one
    // column (the transposed row tile delivers a[k][2q], a[k][2q+1] as one 64-bit shared-memory word), W is
    // duplicated into both halves -- half the FMA issue slots of the scalar loop
    unsigned long long acc2[kEpiRows / 2][NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      unsigned long long bb;
      asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(bj[j]));
#pragma unroll
      for (int q = 0; q < kEpiRows / 2; ++q) acc2[q][j] = bb;
    }
    auto k_step = [&](int k) {
      unsigned long long wv[NC];
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        int c = t + kEpiThreads * j;
        const float w1 = (c < cout) ? s_w[k * cout + c] : 0.f;
        asm("mov.b64 %0, {%1, %1};" : "=l"(wv[j]) : "f"(w1));
      }
      const ulonglong2 a01 = *reinterpret_cast<const ulonglong2*>(s_a + k * kEpiRows);       // rows (0,1), (2,3)
      const ulonglong2 a23 = *reinterpret_cast<const ulonglong2*>(s_a + k * kEpiRows + 4);   // rows (4,5), (6,7)
      const unsigned long long av[kEpiRows / 2] = {a01.x, a01.y, a23.x, a23.y};
#pragma unroll
      for (int q = 0; q < kEpiRows / 2; ++q)
#pragma unroll
        for (int j = 0; j < NC; ++j) asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2[q][j]) : "l"(av[q]), "l"(wv[j]));
    };
    if constexpr (BWD) {
      // compiler's own unrolling: forcing a factor (1 or 4) on this register-capped instantiation cost 1.4-2x
      for (int k = 0; k < cin; ++k) k_step(k);
    } else {
#pragma unroll 4
      for (int k = 0; k < cin; ++k) k_step(k);
    }
    float acc[kEpiRows][NC];
#pragma unroll
    for (int q = 0; q < kEpiRows / 2; ++q)
#pragma unroll
      for (int j = 0; j < NC; ++j)
        asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[2 * q][j]), "=f"(acc[2 * q + 1][j]) : "l"(acc2[q][j]));
    // row norms
    float part[kEpiRows];
#pragma unroll
    for (int rr = 0; rr < kEpiRows; ++rr) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < NC; ++j)
        if (t + kEpiThreads * j < cout) s = fmaf(acc[rr][j], acc[rr][j], s);
      part[rr] = warp_sum(s);
    }
    if (lane == 0) {
#pragma unroll
      for (int rr = 0; rr < kEpiRows; ++rr) s_red[warp * kEpiRows + rr] = part[rr];
    }
    __syncthreads();
    float nrm[kEpiRows], inv[kEpiRows];
#pragma unroll
    for (int rr = 0; rr < kEpiRows; ++rr) {
      float s = 0.f;
#pragma unroll
      for (int wv = 0; wv < kEpiThreads / 32; ++wv) s += s_red[wv * kEpiRows + rr];
      nrm[rr] = sqrtf(s);
      inv[rr] = 1.0f / fmaxf(nrm[rr], kL2NormEps);
    }
    if (!BWD) {
#pragma unroll
      for (int rr = 0; rr < kEpiRows; ++rr) {
        const int64_t r = r0 + rr;
        if (r >= rows) break;
        const bool valid = (vmask >> rr) & 1u;
#pragma unroll
        for (int j = 0; j < NC; ++j) {
          int c = t + kEpiThreads * j;
          if (c < cout) {
            const float v = valid ? act_fwd(act, acc[rr][j] * inv[rr]) : 0.f;
            y[r * ldy + c] = v;
            st1[j] += v;
            st2[j] = fmaf(v, v, st2[j]);
          }
        }
      }
    } else {
      // gm = gy * [y > 0] * mask ; dot_r = sum_c gm * hn      (gy: as given, or formed from the BatchNorm backward)
      float gm[kEpiRows][NC];
#pragma unroll
      for (int rr = 0; rr < kEpiRows; ++rr) {
        const int64_t r = r0 + rr;
        // (backward: per-thread mask evaluation as in round 1a -- routing it through the shared flags made this
        //  register-capped instantiation 1.4-2x slower; the loop order rows-outer / columns-inner is the measured one
        //  too: columns-outer with the BatchNorm coefficients in registers ran 1.7x slower)
        const bool valid = (r < rows) && row_valid(num_nodes, nmax, r);
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < NC; ++j) {
          int c = t + kEpiThreads * j;
          float g = 0.f;
          if (valid && c < cout) {                    // independent loads (no y -> gy dependency chain)
            const float yv = __ldg(y + r * ldy + c);
            float gv = gy[r * ldgy + c];
            if (with_bn) gv = fmaf(s_bn[c], gv, -fmaf(s_bn[cout + c], yv, s_bn[2 * cout + c]));
            g = gv * agrad(yv);
          }
          gm[rr][j] = g;
          s = fmaf(g, acc[rr][j] * inv[rr], s);
        }
        part[rr] = warp_sum(s);
      }
      __syncthreads();                               // everyone has read s_red (norms)
      if (lane == 0) {
#pragma unroll
        for (int rr = 0; rr < kEpiRows; ++rr) s_red[warp * kEpiRows + rr] = part[rr];
      }
      __syncthreads();
#pragma unroll
      for (int rr = 0; rr < kEpiRows; ++rr) {
        const int64_t r = r0 + rr;
        if (r >= rows) break;
        float dot = 0.f;
#pragma unroll
        for (int wv = 0; wv < kEpiThreads / 32; ++wv) dot += s_red[wv * kEpiRows + rr];
#pragma unroll
        for (int j = 0; j < NC; ++j) {
          int c = t + kEpiThreads * j;
          if (c < cout) {
            float hn = acc[rr][j] * inv[rr];
            float d = (nrm[rr] >= kL2NormEps) ? (gm[rr][j] - hn * dot) * inv[rr] : gm[rr][j] * inv[rr];
            dh[r * lddh + c] = d;
          }
        }
      }
    }
  }
  if (!BWD && partial != nullptr) {
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const int c = t + kEpiThreads * j;
      if (c < cout) {
        partial[((int64_t)blockIdx.x * 2 + 0) * cout + c] = (double)st1[j];
        partial[((int64_t)blockIdx.x * 2 + 1) * cout + c] = (double)st2[j];
      }
    }
  }
}

// ------------------------------------------------------------------- wide epilogue backward, one warp per row pair
// The tile kernel above splits a row over the threads of a CTA, so both row reductions (the L2 norm and g . hn) go
// through shared memory and five block barriers per 8 rows: measured 13 800 cycles per 6-row tile at an IPC of ~1
// with the memory system idle (909 us for 1.57 GB).  Here a WARP owns two whole rows: lane l holds the float4
// column chunks l, l + 32, ... (NQ of them), both reductions are shuffles, warps never wait for each other, and the
// accumulators are column pairs in packed f32x2 registers (W pairs come straight out of a 128-bit shared-memory
// load; the broadcast a is duplicated once per k).  The second sweep re-reads y / gy of the two rows (9 KB, L1 / L2)
// instead of holding the masked gradient in 72 more registers.  W is read from shared memory once per row PAIR.
constexpr int kRowPairThreads = 384;      // 12 warps at 168 registers (a few spilled words) beat 8 warps at 255: 655 vs 695 us
constexpr int kRowPairThreadsAlt = 256;   // lab switch CGC_EPI_ROWPAIR_THREADS=256

__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long v;
  asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo), "f"(hi));
  return v;
}
__device__ __forceinline__ float2 unpack2(unsigned long long v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ void ffma2(unsigned long long& acc, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}

// Forward (BWD = false): y = mask * act(h / ||h||), and with partial != null the column sums of y and y^2 for the
// BatchNorm statistics: per lane in registers over the warp's rows, folded warp by warp (fixed order) into one
// [2][cout] partial per CTA, the layout bn_finalize_partials reads.
template <int NQ, bool BWD, int THREADS, bool PADDED>
__global__ void __launch_bounds__(THREADS, 1)
k_epi_rowpair(const float* __restrict__ sa, int64_t ldsa, const float* __restrict__ w,
              const float* __restrict__ bias, const int32_t* __restrict__ num_nodes, int nmax, int64_t rows,
              int cin, int cout, float* y, int64_t ldy, const float* gy, int64_t ldgy,
              float* dh, int64_t lddh, int act, double* __restrict__ partial, const EpiBn bn) {
  extern __shared__ __align__(16) float rp_smem[];
  const ActGrad agrad(act);
  // every per-column vector is padded to cw = cout rounded up to 4 (zero weights / bias / coefficients in the pad
  // columns), so a lane's float4 chunk never straddles a row; pad columns are masked out of loads and stores below
  const int cw = (cout + 3) & ~3;
  const int tail = PADDED ? (cout & 3) : 0;           // valid elements of the last chunk (0 = all four)
  float* s_w = rp_smem;                               // [cin][cw]
  float* s_b = s_w + cin * cw;                        // [cw]
  float* s_bn = s_b + cw;                             // BWD: [3][cw] P, R, Q (BatchNorm backward folded in)
                                                      // forward with statistics: [2][cw] column sums of the CTA
  const bool with_bn = BWD && bn.mean != nullptr;
  float st1[BWD ? 1 : NQ][4], st2[BWD ? 1 : NQ][4];
#pragma unroll
  for (int j = 0; j < (BWD ? 1 : NQ); ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) { st1[j][e] = 0.f; st2[j][e] = 0.f; }
  for (int i = threadIdx.x; i < cin * cw; i += blockDim.x) {
    const int k = i / cw, c = i - k * cw;
    s_w[i] = c < cout ? w[k * cout + c] : 0.f;
  }
  for (int c = threadIdx.x; c < cw; c += blockDim.x) {
    const bool in = c < cout;
    s_b[c] = (bias && in) ? bias[c] : 0.f;
    if (with_bn) {
      const float is = in ? bn.invstd[c] : 0.f, pc = in ? bn.gamma[c] * is : 0.f;
      const float rc = (bn.training && in) ? pc * is * bn.d_gamma[c] * bn.inv_rows : 0.f;
      s_bn[c] = pc;
      s_bn[cw + c] = rc;
      s_bn[2 * cw + c] = (bn.training && in) ? pc * bn.d_beta[c] * bn.inv_rows - rc * bn.mean[c] : 0.f;
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int c4 = cw >> 2;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t npairs = (rows + 1) >> 1;
  // chunk j of this lane; lanes past the last chunk work on a copy of it and are masked out of sums and stores
  int qa[NQ];
  bool live[NQ];
#pragma unroll
  for (int j = 0; j < NQ; ++j) {
    const int q = lane + 32 * j;
    live[j] = q < c4;
    qa[j] = live[j] ? q : c4 - 1;
  }
  // Backward: the y / gy rows of a pair are pulled into L2 one iteration ahead (one 128-byte line per lane and step),
  // so their HBM round trip runs under the shared-memory-bound k loop of the current pair instead of after it.
  auto prefetch_pair = [&](int64_t rp) {
    if (!BWD || rp >= npairs) return;
    const int64_t r0 = 2 * rp;
    const int nrow = (r0 + 1 < rows) ? 2 : 1;
    const int lines = (cout + 31) >> 5;
    for (int i = lane; i < nrow * lines; i += 32) {
      const int rr = i >= lines ? 1 : 0, ch = i - rr * lines;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(y + (r0 + rr) * ldy + ch * 32));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(gy + (r0 + rr) * ldgy + ch * 32));
    }
  };
  prefetch_pair(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  // store / zero one chunk; the last chunk of a row keeps its hands off the columns past cout (they may belong to
  // the next slice of a wider matrix)
  auto store4 = [&](float* row, int q, const float4 v) {
    if (!PADDED || tail == 0 || q != c4 - 1) {
      *reinterpret_cast<float4*>(row + 4 * q) = v;
    } else {
      row[4 * q] = v.x;
      if (tail > 1) row[4 * q + 1] = v.y;
      if (tail > 2) row[4 * q + 2] = v.z;
    }
  };
  auto mask4 = [&](float4& v, int q) {                // zero what was read past cout (garbage, possibly NaN)
    if (PADDED && tail != 0 && q == c4 - 1) {
      if (tail < 2) v.y = 0.f;
      if (tail < 3) v.z = 0.f;
      v.w = 0.f;
    }
  };
  for (int64_t rp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; rp < npairs; rp += warps) {
    const int64_t r0 = 2 * rp, r1 = r0 + 1;
    const bool has1 = r1 < rows;
    prefetch_pair(rp + warps);
    const bool v0 = row_valid(num_nodes, nmax, r0), v1 = has1 && row_valid(num_nodes, nmax, r1);
    if (!v0 && !v1) {                                 // padding rows: zero output / gradient
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      float* o = BWD ? dh : y;
      const int64_t ldo = BWD ? lddh : ldy;
#pragma unroll
      for (int j = 0; j < NQ; ++j)
        if (live[j]) {
          store4(o + r0 * ldo, qa[j], z);
          if (has1) store4(o + r1 * ldo, qa[j], z);
        }
      continue;
    }
    // aggregate rows: lane k holds a[k] (and a[k + 32])
    const float* a0p = sa + r0 * ldsa;
    const float* a1p = sa + (has1 ? r1 : r0) * ldsa;
    const float a0lo = lane < cin ? __ldg(a0p + lane) : 0.f, a0hi = lane + 32 < cin ? __ldg(a0p + lane + 32) : 0.f;
    const float a1lo = lane < cin ? __ldg(a1p + lane) : 0.f, a1hi = lane + 32 < cin ? __ldg(a1p + lane + 32) : 0.f;
    unsigned long long acc0[NQ][2], acc1[NQ][2];      // h of rows r0 / r1: column pairs (4q, 4q+1), (4q+2, 4q+3)
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
      const ulonglong2 b = *reinterpret_cast<const ulonglong2*>(s_b + 4 * qa[j]);
      acc0[j][0] = b.x; acc0[j][1] = b.y;
      acc1[j][0] = b.x; acc1[j][1] = b.y;
    }
    auto k_step = [&](int k, float a0, float a1) {
      const unsigned long long a0d = pack2(a0, a0), a1d = pack2(a1, a1);
      const float* wrow = s_w + k * cw;
#pragma unroll
      for (int j = 0; j < NQ; ++j) {
        const ulonglong2 wv = *reinterpret_cast<const ulonglong2*>(wrow + 4 * qa[j]);
        ffma2(acc0[j][0], a0d, wv.x);
        ffma2(acc0[j][1], a0d, wv.y);
        ffma2(acc1[j][0], a1d, wv.x);
        ffma2(acc1[j][1], a1d, wv.y);
      }
    };
    const int klo = cin < 32 ? cin : 32;
#pragma unroll 2
    for (int k = 0; k < klo; ++k) k_step(k, __shfl_sync(0xffffffffu, a0lo, k), __shfl_sync(0xffffffffu, a1lo, k));
    for (int k = 32; k < cin; ++k) k_step(k, __shfl_sync(0xffffffffu, a0hi, k - 32), __shfl_sync(0xffffffffu, a1hi, k - 32));
    // row norms
    float sq0 = 0.f, sq1 = 0.f;
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
      if (live[j]) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float2 p0 = unpack2(acc0[j][h]), p1 = unpack2(acc1[j][h]);
          sq0 = fmaf(p0.x, p0.x, fmaf(p0.y, p0.y, sq0));
          sq1 = fmaf(p1.x, p1.x, fmaf(p1.y, p1.y, sq1));
        }
      }
    }
    const float nrm0 = sqrtf(warp_sum(sq0)), nrm1 = sqrtf(warp_sum(sq1));
    const float inv0 = 1.0f / fmaxf(nrm0, kL2NormEps), inv1 = 1.0f / fmaxf(nrm1, kL2NormEps);
    if constexpr (!BWD) {
#pragma unroll
      for (int j = 0; j < NQ; ++j) {
        if (live[j]) {
          const float2 h00 = unpack2(acc0[j][0]), h01 = unpack2(acc0[j][1]);
          const float2 h10 = unpack2(acc1[j][0]), h11 = unpack2(acc1[j][1]);
          float4 oa = make_float4(0.f, 0.f, 0.f, 0.f), ob = oa;
          if (v0) oa = make_float4(act_fwd(act, h00.x * inv0), act_fwd(act, h00.y * inv0), act_fwd(act, h01.x * inv0),
                                   act_fwd(act, h01.y * inv0));
          if (v1) ob = make_float4(act_fwd(act, h10.x * inv1), act_fwd(act, h10.y * inv1), act_fwd(act, h11.x * inv1),
                                   act_fwd(act, h11.y * inv1));
          store4(y + r0 * ldy, qa[j], oa);
          if (has1) store4(y + r1 * ldy, qa[j], ob);
          st1[j][0] += oa.x + ob.x; st1[j][1] += oa.y + ob.y; st1[j][2] += oa.z + ob.z; st1[j][3] += oa.w + ob.w;
          st2[j][0] = fmaf(oa.x, oa.x, fmaf(ob.x, ob.x, st2[j][0]));
          st2[j][1] = fmaf(oa.y, oa.y, fmaf(ob.y, ob.y, st2[j][1]));
          st2[j][2] = fmaf(oa.z, oa.z, fmaf(ob.z, ob.z, st2[j][2]));
          st2[j][3] = fmaf(oa.w, oa.w, fmaf(ob.w, ob.w, st2[j][3]));
        }
      }
      continue;
    }
    // masked gradient of one chunk: g = gy * act'(y)   (gy formed from the BatchNorm backward when folded in)
    auto grad4 = [&](const float4 yv, float4 gv, int q) {
      if (with_bn) {
        const float4 pc = *reinterpret_cast<const float4*>(s_bn + 4 * q);
        const float4 rc = *reinterpret_cast<const float4*>(s_bn + cw + 4 * q);
        const float4 qc = *reinterpret_cast<const float4*>(s_bn + 2 * cw + 4 * q);
        gv.x = fmaf(pc.x, gv.x, -fmaf(rc.x, yv.x, qc.x));
        gv.y = fmaf(pc.y, gv.y, -fmaf(rc.y, yv.y, qc.y));
        gv.z = fmaf(pc.z, gv.z, -fmaf(rc.z, yv.z, qc.z));
        gv.w = fmaf(pc.w, gv.w, -fmaf(rc.w, yv.w, qc.w));
      }
      return make_float4(gv.x * agrad(yv.x), gv.y * agrad(yv.y), gv.z * agrad(yv.z), gv.w * agrad(yv.w));
    };
    const float* y0 = y + r0 * ldy;
    const float* y1 = y + (has1 ? r1 : r0) * ldy;
    const float* g0 = gy + r0 * ldgy;
    const float* g1 = gy + (has1 ? r1 : r0) * ldgy;
    // sweep 1: dot_r = sum_c g * h   (scaled by inv below)
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
      const float4 ya = __ldg(reinterpret_cast<const float4*>(y0) + qa[j]);
      const float4 yb = __ldg(reinterpret_cast<const float4*>(y1) + qa[j]);
      const float4 ga = *(reinterpret_cast<const float4*>(g0) + qa[j]);
      const float4 gb = *(reinterpret_cast<const float4*>(g1) + qa[j]);
      if (live[j]) {
        float4 ma = grad4(ya, ga, qa[j]), mb = grad4(yb, gb, qa[j]);
        mask4(ma, qa[j]);
        mask4(mb, qa[j]);
        const float2 h00 = unpack2(acc0[j][0]), h01 = unpack2(acc0[j][1]);
        const float2 h10 = unpack2(acc1[j][0]), h11 = unpack2(acc1[j][1]);
        d0 = fmaf(ma.x, h00.x, fmaf(ma.y, h00.y, fmaf(ma.z, h01.x, fmaf(ma.w, h01.y, d0))));
        d1 = fmaf(mb.x, h10.x, fmaf(mb.y, h10.y, fmaf(mb.z, h11.x, fmaf(mb.w, h11.y, d1))));
      }
    }
    const float dot0 = v0 ? warp_sum(d0) * inv0 : 0.f, dot1 = v1 ? warp_sum(d1) * inv1 : 0.f;
    // sweep 2: dh = (g - hn dot) / ||h||   (g / eps where the norm was clamped); y / gy again, from L1 / L2
    const bool big0 = nrm0 >= kL2NormEps, big1 = nrm1 >= kL2NormEps;
    const float m0 = v0 ? 1.f : 0.f, m1 = v1 ? 1.f : 0.f;
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
      const float4 ya = __ldg(reinterpret_cast<const float4*>(y0) + qa[j]);
      const float4 yb = __ldg(reinterpret_cast<const float4*>(y1) + qa[j]);
      const float4 ga = *(reinterpret_cast<const float4*>(g0) + qa[j]);
      const float4 gb = *(reinterpret_cast<const float4*>(g1) + qa[j]);
      if (live[j]) {
        float4 ma = grad4(ya, ga, qa[j]), mb = grad4(yb, gb, qa[j]);
        mask4(ma, qa[j]);
        mask4(mb, qa[j]);
        ma.x *= m0; ma.y *= m0; ma.z *= m0; ma.w *= m0;
        mb.x *= m1; mb.y *= m1; mb.z *= m1; mb.w *= m1;
        const float2 h00 = unpack2(acc0[j][0]), h01 = unpack2(acc0[j][1]);
        const float2 h10 = unpack2(acc1[j][0]), h11 = unpack2(acc1[j][1]);
        const float s0 = big0 ? inv0 * dot0 : 0.f, s1 = big1 ? inv1 * dot1 : 0.f;     // hn dot = h (inv dot)
        float4 oa, ob;
        oa.x = (ma.x - h00.x * s0) * inv0; oa.y = (ma.y - h00.y * s0) * inv0;
        oa.z = (ma.z - h01.x * s0) * inv0; oa.w = (ma.w - h01.y * s0) * inv0;
        ob.x = (mb.x - h10.x * s1) * inv1; ob.y = (mb.y - h10.y * s1) * inv1;
        ob.z = (mb.z - h11.x * s1) * inv1; ob.w = (mb.w - h11.y * s1) * inv1;
        store4(dh + r0 * lddh, qa[j], oa);
        if (has1) store4(dh + r1 * lddh, qa[j], ob);
      }
    }
  }
  if constexpr (!BWD) {
    if (partial != nullptr) {
      const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
      for (int wv = 0; wv < nwarps; ++wv) {            // warp 0 writes, the others add in warp order
        if (warp == wv) {
#pragma unroll
          for (int j = 0; j < NQ; ++j) {
            if (live[j]) {
              float4* p1 = reinterpret_cast<float4*>(s_bn + 4 * qa[j]);
              float4* p2 = reinterpret_cast<float4*>(s_bn + cw + 4 * qa[j]);
              float4 a = make_float4(st1[j][0], st1[j][1], st1[j][2], st1[j][3]);
              float4 b = make_float4(st2[j][0], st2[j][1], st2[j][2], st2[j][3]);
              if (wv != 0) {
                const float4 pa = *p1, pb = *p2;
                a.x += pa.x; a.y += pa.y; a.z += pa.z; a.w += pa.w;
                b.x += pb.x; b.y += pb.y; b.z += pb.z; b.w += pb.w;
              }
              *p1 = a;
              *p2 = b;
            }
          }
        }
        __syncthreads();
      }
      for (int c = threadIdx.x; c < cout; c += blockDim.x) {
        partial[((int64_t)blockIdx.x * 2 + 0) * cout + c] = (double)s_bn[c];
        partial[((int64_t)blockIdx.x * 2 + 1) * cout + c] = (double)s_bn[cw + c];
      }
    }
  }
}

static size_t rowpair_smem_bytes(int cin, int cout, int extra_vectors) {
  const size_t cw = ((size_t)cout + 3) & ~(size_t)3;
  return sizeof(float) * ((size_t)cin * cw + cw + (size_t)extra_vectors * cw);
}

template <bool BWD>
static int launch_epilogue(const float* sa, int64_t ldsa, const float* w, const float* b,
                           const int32_t* num_nodes, int nmax, int64_t rows, int cin, int cout, float* y,
                           int64_t ldy, const float* gy, int64_t ldgy, float* dh, int64_t lddh, int act,
                           cudaStream_t st, double* partial = nullptr, int64_t work_bytes = 0, int* blocks_out = nullptr,
                           const EpiBn* bn_in = nullptr) {
  EpiBn bn;
  memset(&bn, 0, sizeof(bn));
  if (bn_in != nullptr) bn = *bn_in;
  CGC_REQUIRE(act >= CGC_ACT_RELU && act <= CGC_ACT_LEAKY_RELU, CGC_ERR_ARG, "sage epilogue: unknown activation %d", act);
  if (cout <= 32) {
    CGC_REQUIRE(bn_in == nullptr, CGC_ERR_UNSUPPORTED, "sage epilogue: BatchNorm folding needs cout > 32");
    CGC_REQUIRE(cin <= 64, CGC_ERR_UNSUPPORTED, "sage epilogue: cin=%d > 64", cin);
    size_t smem = sizeof(float) * (size_t)(cin * cout + cout);
    int64_t warps_needed = rows;
    int64_t blocks = ceil_div(warps_needed, 8);
    if (blocks > kNumSMs * 32) blocks = kNumSMs * 32;   // latency-bound per row: keep many warps in flight
    k_epi_narrow<BWD><<<(unsigned)blocks, 256, smem, st>>>(sa, ldsa, w, b, num_nodes, nmax, rows, cin, cout, y, ldy,
                                                           gy, ldgy, dh, lddh, act);
    CGC_CHECK_LAUNCH("k_epi_narrow");
    return 0;
  }
  const int nc = (int)ceil_div(cout, kEpiThreads);
  CGC_REQUIRE(nc <= 5, CGC_ERR_UNSUPPORTED, "sage epilogue: cout=%d > 1280", cout);
  CGC_REQUIRE(bn_in == nullptr || BWD, CGC_ERR_ARG, "sage epilogue: BatchNorm folding is a backward option");
  {
    // wide layers with 16-byte aligned rows: one warp per row pair
    const int nq = (int)ceil_div((cout + 3) / 4, 32);
    static const bool rowpair_off = lab_env("CGC_EPI_ROWPAIR", 1) == 0;   // A/B switch (CGC_LAB builds only)
    const int extra = BWD ? (bn_in != nullptr ? 3 : 0) : (partial != nullptr ? 2 : 0);
    float* o = BWD ? dh : y;
    const int64_t ldo = BWD ? lddh : ldy;
    // (measured at 20 x 1140, 115 200 rows: backward 653 us vs 1076 us for the tile kernel -- 909 us when that was fed
    //  by a two-stage cp.async.bulk ring, a variant since removed; forward 369 us vs 333 us tile kernel -- W is read from shared memory once per row PAIR here, once per 8 rows there, and the
    //  forward has no global loads to overlap: BWD only unless the lab switch is 2)
    static const bool rowpair_fwd = lab_env("CGC_EPI_ROWPAIR", 1) == 2;
    // any width from 64 columns up: vectors are padded to a multiple of 4 in shared memory, pad columns masked
    // (cout = 114, stage 2: 155 us in the tile kernel, where 142 of 256 threads own no column)
    if ((BWD || rowpair_fwd) && !rowpair_off && cout >= 64 && nq <= 10 && cin <= 64 &&
        ldy % 4 == 0 && aligned16(y) &&
        ldo % 4 == 0 && aligned16(o) && (!BWD || (ldgy % 4 == 0 && aligned16(gy))) &&
        rowpair_smem_bytes(cin, cout, extra) <= 200 * 1024) {
      const size_t smem = rowpair_smem_bytes(cin, cout, extra);
      static const int threads =
          lab_env("CGC_EPI_ROWPAIR_THREADS", kRowPairThreads) == kRowPairThreadsAlt ? kRowPairThreadsAlt : kRowPairThreads;
      int64_t blocks = ceil_div(ceil_div(rows, 2), threads / 32);
      if (blocks > kNumSMs) blocks = kNumSMs;
      if (partial != nullptr) {
        CGC_REQUIRE(blocks * 2 * cout * 8 <= work_bytes, CGC_ERR_WORKSPACE, "sage epilogue: workspace too small");
        *blocks_out = (int)blocks;
      }
#define CGC_ROWPAIR_LAUNCH(N, T)                                                                               \
  {                                                                                                            \
    auto kern = (cout & 3) ? k_epi_rowpair<N, BWD, T, true> : k_epi_rowpair<N, BWD, T, false>;                 \
    CGC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));              \
    kern<<<(unsigned)blocks, T, smem, st>>>(sa, ldsa, w, b, num_nodes, nmax, rows, cin, cout, y, ldy, gy, ldgy, \
                                            dh, lddh, act, partial, bn);                                       \
  }
#define CGC_ROWPAIR_CASE(N)                                                                                    \
  case N:                                                                                                      \
    if (threads == kRowPairThreads) CGC_ROWPAIR_LAUNCH(N, kRowPairThreads) else CGC_ROWPAIR_LAUNCH(N, kRowPairThreadsAlt) \
    break;
      switch (nq) {
        CGC_ROWPAIR_CASE(1)
        CGC_ROWPAIR_CASE(2)
        CGC_ROWPAIR_CASE(3)
        CGC_ROWPAIR_CASE(4)
        CGC_ROWPAIR_CASE(5)
        CGC_ROWPAIR_CASE(6)
        CGC_ROWPAIR_CASE(7)
        CGC_ROWPAIR_CASE(8)
        CGC_ROWPAIR_CASE(9)
        CGC_ROWPAIR_CASE(10)
      }
#undef CGC_ROWPAIR_CASE
#undef CGC_ROWPAIR_LAUNCH
      CGC_CHECK_LAUNCH("k_epi_rowpair");
      return 0;
    }
  }
  size_t smem = sizeof(float) * (size_t)(cin * cout + cin * kEpiRows + (kEpiThreads / 32) * kEpiRows + kEpiRows +
                                         (bn_in != nullptr ? 3 * cout : 0));
  CGC_REQUIRE(smem <= 200 * 1024, CGC_ERR_UNSUPPORTED, "sage epilogue: W of %dx%d does not fit in shared memory",
              cin, cout);
  int64_t tiles = ceil_div(rows, kEpiRows);
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  int64_t blocks = tiles < (int64_t)kNumSMs * per_sm ? tiles : (int64_t)kNumSMs * per_sm;
  if (partial != nullptr) {
    CGC_REQUIRE(blocks * 2 * cout * 8 <= work_bytes, CGC_ERR_WORKSPACE, "sage epilogue: workspace too small");
    *blocks_out = (int)blocks;
  }
#define CGC_EPI_CASE(N)                                                                                       \
  case N: {                                                                                                   \
    auto kern = k_epi_wide<N, BWD>;                                                                           \
    CGC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));             \
    kern<<<(unsigned)blocks, kEpiThreads, smem, st>>>(sa, ldsa, w, b, num_nodes, nmax, rows, cin, cout, y, ldy, \
                                                      gy, ldgy, dh, lddh, act, partial, bn);                  \
  } break;
  switch (nc) {
    CGC_EPI_CASE(1)
    CGC_EPI_CASE(2)
    CGC_EPI_CASE(3)
    CGC_EPI_CASE(4)
    CGC_EPI_CASE(5)
  }
#undef CGC_EPI_CASE
  CGC_CHECK_LAUNCH("k_epi_wide");
  return 0;
}

// =================================================================== softmax over clusters
// 16-byte global loads in program order (asm volatile: the compiler may not sink them behind their first use),
// so a lane has all its requests in flight before it waits for the first one
__device__ __forceinline__ float4 ldg_nc_f4_ordered(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// vectorised variant: c % 4 == 0, 16-byte aligned rows; lane owns float4 chunks lane, lane + 32, ... (NV4 of them)
template <int NV4>
__global__ void __launch_bounds__(256)
k_softmax_fwd_v4(const float* __restrict__ z, int64_t rows, int c, const int32_t* __restrict__ num_nodes, int nmax,
                 float* __restrict__ s, int64_t lds, float* __restrict__ s_full, int32_t* __restrict__ arg) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= rows) return;
  const float* zr = z + r * (int64_t)c;
  const int c4 = c >> 2;
  float4 v[NV4];
#pragma unroll
  for (int j = 0; j < NV4; ++j) {
    const int q = lane + 32 * j;
    v[j] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    if (q < c4) v[j] = ldg_nc_f4_ordered(zr + 4 * q);
  }
  float m = -INFINITY;
  int mi = 0x7fffffff;
#pragma unroll
  for (int j = 0; j < NV4; ++j) {
    const int i = 4 * (lane + 32 * j);
    if (v[j].x > m) { m = v[j].x; mi = i; }
    if (v[j].y > m) { m = v[j].y; mi = i + 1; }
    if (v[j].z > m) { m = v[j].z; mi = i + 2; }
    if (v[j].w > m) { m = v[j].w; mi = i + 3; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float om = __shfl_xor_sync(0xffffffffu, m, o);
    int oi = __shfl_xor_sync(0xffffffffu, mi, o);
    if (om > m || (om == m && oi < mi)) { m = om; mi = oi; }
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < NV4; ++j) {
    const bool on = lane + 32 * j < c4;
    v[j].x = on ? expf(v[j].x - m) : 0.f; v[j].y = on ? expf(v[j].y - m) : 0.f;
    v[j].z = on ? expf(v[j].z - m) : 0.f; v[j].w = on ? expf(v[j].w - m) : 0.f;
    sum += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  const bool valid = row_valid(num_nodes, nmax, r);
  float* sr = s + r * lds;
  float* fr = s_full ? s_full + r * (int64_t)c : nullptr;
#pragma unroll
  for (int j = 0; j < NV4; ++j) {
    const int q = lane + 32 * j;
    if (q < c4) {
      const float4 pv = make_float4(v[j].x * inv, v[j].y * inv, v[j].z * inv, v[j].w * inv);
      *reinterpret_cast<float4*>(sr + 4 * q) = valid ? pv : make_float4(0.f, 0.f, 0.f, 0.f);
      if (fr) *reinterpret_cast<float4*>(fr + 4 * q) = pv;
    }
  }
  if (arg && lane == 0) arg[r] = mi;
}

// one warp per row, row cached in registers (NREG values per lane)
template <int NREG>
__global__ void __launch_bounds__(256)
k_softmax_fwd(const float* __restrict__ z, int64_t rows, int c, const int32_t* __restrict__ num_nodes, int nmax,
              float* __restrict__ s, int64_t lds, float* __restrict__ s_full, int32_t* __restrict__ arg) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= rows) return;
  const float* zr = z + r * (int64_t)c;
  float v[NREG];
  // all loads first (NREG independent requests in flight per lane), then the running max: a compare inside the
  // load loop chains every load behind the previous one's DRAM latency (measured: 509 us -> see DESIGN.md)
#pragma unroll
  for (int j = 0; j < NREG; ++j) {
    int i = lane + 32 * j;
    v[j] = (i < c) ? __ldg(zr + i) : -INFINITY;
  }
  float m = -INFINITY;
  int mi = 0x7fffffff;
#pragma unroll
  for (int j = 0; j < NREG; ++j) {
    if (v[j] > m) { m = v[j]; mi = lane + 32 * j; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float om = __shfl_xor_sync(0xffffffffu, m, o);
    int oi = __shfl_xor_sync(0xffffffffu, mi, o);
    if (om > m || (om == m && oi < mi)) { m = om; mi = oi; }
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < NREG; ++j) {
    int i = lane + 32 * j;
    v[j] = (i < c) ? expf(v[j] - m) : 0.f;
    sum += v[j];
  }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  const bool valid = row_valid(num_nodes, nmax, r);
  float* sr = s + r * lds;
  float* fr = s_full ? s_full + r * (int64_t)c : nullptr;
#pragma unroll
  for (int j = 0; j < NREG; ++j) {
    int i = lane + 32 * j;
    if (i < c) {
      float p = v[j] * inv;
      sr[i] = valid ? p : 0.f;
      if (fr) fr[i] = p;
    }
  }
  if (arg && lane == 0) arg[r] = mi;
}

// one warp per row; both rows cached in registers (NREG values per lane), every load issued before the first use
template <int NREG>
__global__ void __launch_bounds__(256)
k_softmax_bwd(const float* __restrict__ gs, const float* __restrict__ s, int64_t lds, int64_t rows, int c,
              float* __restrict__ gz, int64_t ldgz) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= rows) return;
  const float* g = gs + r * lds;
  const float* p = s + r * lds;
  float gv[NREG], pv[NREG];
#pragma unroll
  for (int j = 0; j < NREG; ++j) {
    const int i = lane + 32 * j;
    gv[j] = (i < c) ? __ldg(g + i) : 0.f;
    pv[j] = (i < c) ? __ldg(p + i) : 0.f;
  }
  float dot = 0.f;
#pragma unroll
  for (int j = 0; j < NREG; ++j) dot = fmaf(gv[j], pv[j], dot);
  dot = warp_sum(dot);
  float* o = gz + r * ldgz;
#pragma unroll
  for (int j = 0; j < NREG; ++j) {
    const int i = lane + 32 * j;
    if (i < c) o[i] = pv[j] * (gv[j] - dot);
  }
}

// =================================================================== small row helpers
__global__ void __launch_bounds__(256)
k_scale_rows_bwd(const float* __restrict__ dsa, const float* __restrict__ sa, const float* __restrict__ rs,
                 int64_t rows, int c, float* __restrict__ ga, int64_t ldga, float* __restrict__ d_rs) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= rows) return;
  const float sc = rs[r];
  float dot = 0.f;
  for (int i = lane; i < c; i += 32) {
    float d = dsa[r * c + i];
    dot = fmaf(d, sa[r * c + i], dot);
    ga[r * ldga + i] = d * sc;
  }
  dot = warp_sum(dot);
  if (lane == 0 && d_rs) d_rs[r] += dot / sc;
}

__global__ void __launch_bounds__(256)
k_row_inv_clamped_sum(const float* __restrict__ adj, int64_t rows, int n, float* __restrict__ rs,
                      float* __restrict__ rsum) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= rows) return;
  const float* a = adj + r * (int64_t)n;
  float s = 0.f;
  for (int i = lane; i < n; i += 32) s += __ldg(a + i);
  s = warp_sum(s);
  if (lane == 0) {
    rsum[r] = s;
    rs[r] = 1.0f / fmaxf(s, 1.0f);
  }
}

__global__ void __launch_bounds__(256)
k_add_rs_grad(float* __restrict__ d_adj, const float* __restrict__ rsum, const float* __restrict__ d_rs,
              int64_t rows, int n) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= rows) return;
  const float s = rsum[r];
  if (!(s >= 1.0f)) return;
  const float v = -d_rs[r] / (s * s);
  float* a = d_adj + r * (int64_t)n;
  for (int i = lane; i < n; i += 32) a[i] += v;
}

__global__ void __launch_bounds__(256)
k_dense_renorm_fwd(const float* __restrict__ adj, int64_t rows, int n, float p, float omp,
                   const int32_t* __restrict__ num_nodes, float* __restrict__ out, float* __restrict__ rho) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= rows) return;
  const int diag = (int)(r % n);
  const float* a = adj + r * (int64_t)n;
  float s = 0.f;
  for (int i = lane; i < n; i += 32) s += (i == diag) ? 0.f : __ldg(a + i);
  s = warp_sum(s) + kRenormEps;
  const bool valid = row_valid(num_nodes, n, r);
  float* o = out + r * (int64_t)n;
  for (int i = lane; i < n; i += 32) {
    float v = (i == diag) ? p : (a[i] / s) * omp;
    o[i] = valid ? v : 0.f;
  }
  if (lane == 0) rho[r] = s;
}

__global__ void __launch_bounds__(256)
k_dense_renorm_bwd(const float* __restrict__ g, const float* __restrict__ out, const float* __restrict__ rho,
                   int64_t rows, int n, float omp, const int32_t* __restrict__ num_nodes,
                   float* __restrict__ g_adj) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= rows) return;
  const int diag = (int)(r % n);
  const bool valid = row_valid(num_nodes, n, r);
  const float* gr = g + r * (int64_t)n;
  const float* nr = out + r * (int64_t)n;
  float* o = g_adj + r * (int64_t)n;
  if (!valid) {
    for (int i = lane; i < n; i += 32) o[i] = 0.f;
    return;
  }
  float dot = 0.f;
  for (int i = lane; i < n; i += 32)
    if (i != diag) dot = fmaf(__ldg(gr + i), __ldg(nr + i), dot);
  dot = warp_sum(dot) / omp;
  const float sc = omp / rho[r];
  for (int i = lane; i < n; i += 32) o[i] = (i == diag) ? 0.f : sc * (gr[i] - dot);
}

// elementwise activation on a [rows, c] matrix with leading dimensions (GIN's MLP and post-conv activation)
template <bool BWD>
__global__ void __launch_bounds__(256) k_act(const float* __restrict__ a, int64_t lda, const float* __restrict__ y,
                                             int64_t ldy, int64_t rows, int c, int act, float* __restrict__ o,
                                             int64_t ldo) {
  const int64_t n = rows * c;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / c;
    const int k = (int)(i - r * c);
    if (!BWD) o[r * ldo + k] = act_fwd(act, a[r * lda + k]);
    else o[r * ldo + k] = a[r * lda + k] * act_grad(act, __ldg(y + r * ldy + k));
  }
}

}  // namespace cgc

using namespace cgc;

extern "C" int cgc_act_fwd(const float* x, int64_t ldx, int64_t rows, int32_t c, int act, float* y, int64_t ldy,
                           void* stream) {
  CGC_REQUIRE(x && y && rows > 0 && c > 0 && ldx >= c && ldy >= c, CGC_ERR_ARG, "cgc_act_fwd: bad args");
  CGC_REQUIRE(act >= CGC_ACT_RELU && act <= CGC_ACT_LEAKY_RELU, CGC_ERR_ARG, "cgc_act_fwd: unknown activation %d", act);
  int64_t blocks = ceil_div(rows * c, 256 * 4);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  k_act<false><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(x, ldx, nullptr, 0, rows, c, act, y, ldy);
  CGC_CHECK_LAUNCH("k_act");
  return 0;
}

extern "C" int cgc_act_bwd(const float* gy, int64_t ldgy, const float* y, int64_t ldy, int64_t rows, int32_t c,
                           int act, float* gx, int64_t ldgx, void* stream) {
  CGC_REQUIRE(gy && y && gx && rows > 0 && c > 0 && ldgy >= c && ldy >= c && ldgx >= c, CGC_ERR_ARG,
              "cgc_act_bwd: bad args");
  CGC_REQUIRE(act >= CGC_ACT_RELU && act <= CGC_ACT_LEAKY_RELU, CGC_ERR_ARG, "cgc_act_bwd: unknown activation %d", act);
  int64_t blocks = ceil_div(rows * c, 256 * 4);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  k_act<true><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(gy, ldgy, y, ldy, rows, c, act, gx, ldgx);
  CGC_CHECK_LAUNCH("k_act");
  return 0;
}

extern "C" int cgc_sage_epilogue_fwd(const float* sa, int64_t ldsa, const float* w, const float* b,
                                     const int32_t* num_nodes, int32_t nmax, int64_t rows, int32_t cin,
                                     int32_t cout, int act, float* y, int64_t ldy, void* stream) {
  CGC_REQUIRE(sa && w && y && rows > 0 && cin > 0 && cout > 0 && ldsa >= cin && ldy >= cout, CGC_ERR_ARG,
              "cgc_sage_epilogue_fwd: bad args");
  CGC_REQUIRE(num_nodes == nullptr || nmax > 0, CGC_ERR_ARG, "cgc_sage_epilogue_fwd: mask without nmax");
  return launch_epilogue<false>(sa, ldsa, w, b, num_nodes, nmax, rows, cin, cout, y, ldy, nullptr, 0, nullptr, 0,
                                act, as_stream(stream));
}

namespace cgc {
int bn_finalize_partials(const double* partial, int blocks, int c, int64_t rows, float* running_mean,
                         float* running_var, float momentum, float eps, float* mean, float* invstd, cudaStream_t st);
}

extern "C" int cgc_sage_layer_fwd_wide(const float* sa, int64_t ldsa, const float* w, const float* b,
                                       const int32_t* num_nodes, int32_t nmax, int64_t rows, int32_t cin,
                                       int32_t cout, int act, float* y, int64_t ldy, float* running_mean,
                                       float* running_var, float momentum, float eps, float* mean, float* invstd,
                                       void* work, int64_t work_bytes, void* stream) {
  CGC_REQUIRE(sa && w && y && mean && invstd && work && rows > 0 && cin > 0 && cout > 0 && ldsa >= cin && ldy >= cout,
              CGC_ERR_ARG, "cgc_sage_layer_fwd_wide: bad args");
  CGC_REQUIRE(cout > 32 && cout <= 1280, CGC_ERR_UNSUPPORTED,
              "cgc_sage_layer_fwd_wide: needs 32 < cout <= 1280 (got %d); narrow layers take cgc_sage_layer_fwd_narrow", cout);
  CGC_REQUIRE(num_nodes == nullptr || nmax > 0, CGC_ERR_ARG, "cgc_sage_layer_fwd_wide: mask without nmax");
  int blocks = 0;
  cudaStream_t st = as_stream(stream);
  int rc = launch_epilogue<false>(sa, ldsa, w, b, num_nodes, nmax, rows, cin, cout, y, ldy, nullptr, 0, nullptr, 0, act,
                                  st, reinterpret_cast<double*>(work), work_bytes, &blocks);
  if (rc != 0) return rc;
  return bn_finalize_partials(reinterpret_cast<const double*>(work), blocks, cout, rows, running_mean, running_var,
                              momentum, eps, mean, invstd, st);
}

extern "C" int cgc_sage_dh(const float* gy, int64_t ldgy, const float* sa, int64_t ldsa, const float* w,
                           const float* b, const float* y, int64_t ldy, const int32_t* num_nodes, int32_t nmax,
                           int64_t rows, int32_t cin, int32_t cout, int act, float* dh, int64_t lddh, void* stream) {
  CGC_REQUIRE(gy && sa && w && y && dh && rows > 0 && cin > 0 && cout > 0, CGC_ERR_ARG, "cgc_sage_dh: bad args");
  CGC_REQUIRE(num_nodes == nullptr || nmax > 0, CGC_ERR_ARG, "cgc_sage_dh: mask without nmax");
  return launch_epilogue<true>(sa, ldsa, w, b, num_nodes, nmax, rows, cin, cout, const_cast<float*>(y), ldy, gy,
                               ldgy, dh, lddh, act, as_stream(stream));
}

extern "C" int cgc_sage_dh_bn(const float* gx, int64_t ldgx, const float* sa, int64_t ldsa, const float* w,
                              const float* b, const float* y, int64_t ldy, const float* mean, const float* invstd,
                              const float* gamma, const float* d_gamma, const float* d_beta, int training,
                              const int32_t* num_nodes, int32_t nmax, int64_t rows, int32_t cin, int32_t cout, int act,
                              float* dh, int64_t lddh, void* stream) {
  CGC_REQUIRE(gx && sa && w && y && dh && mean && invstd && gamma && d_gamma && d_beta && rows > 0 && cin > 0 && cout > 0,
              CGC_ERR_ARG, "cgc_sage_dh_bn: bad args");
  CGC_REQUIRE(cout > 32 && cout <= 1280, CGC_ERR_UNSUPPORTED, "cgc_sage_dh_bn: needs 32 < cout <= 1280 (got %d)", cout);
  CGC_REQUIRE(num_nodes == nullptr || nmax > 0, CGC_ERR_ARG, "cgc_sage_dh_bn: mask without nmax");
  EpiBn bn;
  bn.mean = mean; bn.invstd = invstd; bn.gamma = gamma; bn.d_gamma = d_gamma; bn.d_beta = d_beta;
  bn.inv_rows = 1.0f / (float)rows;
  bn.training = training;
  return launch_epilogue<true>(sa, ldsa, w, b, num_nodes, nmax, rows, cin, cout, const_cast<float*>(y), ldy, gx, ldgx,
                               dh, lddh, act, as_stream(stream), nullptr, 0, nullptr, &bn);
}

extern "C" int cgc_scale_rows_bwd(const float* dsa, const float* sa, const float* rs, int64_t rows, int32_t c,
                                  float* ga, int64_t ldga, float* d_rs, void* stream) {
  CGC_REQUIRE(dsa && sa && rs && ga && rows > 0 && c > 0 && ldga >= c, CGC_ERR_ARG,
              "cgc_scale_rows_bwd: bad args");
  k_scale_rows_bwd<<<(unsigned)ceil_div(rows, 8), 256, 0, as_stream(stream)>>>(dsa, sa, rs, rows, c, ga, ldga,
                                                                               d_rs);
  CGC_CHECK_LAUNCH("k_scale_rows_bwd");
  return 0;
}

extern "C" int cgc_softmax_fwd(const float* z, int64_t rows, int32_t c, const int32_t* num_nodes, int32_t nmax,
                               float* s, int64_t lds, float* s_full, int32_t* arg, void* stream) {
  CGC_REQUIRE(z && s && rows > 0 && c > 0 && lds >= c, CGC_ERR_ARG, "cgc_softmax_fwd: bad args");
  CGC_REQUIRE(c <= 32 * 40, CGC_ERR_UNSUPPORTED, "cgc_softmax_fwd: %d clusters > 1280", c);
  cudaStream_t st = as_stream(stream);
  unsigned grid = (unsigned)ceil_div(rows, 8);
  const bool vec = c % 4 == 0 && lds % 4 == 0 && aligned16(z) && aligned16(s) && (s_full == nullptr || aligned16(s_full));
  if (vec && c > 128) {
    if (c <= 512) k_softmax_fwd_v4<4><<<grid, 256, 0, st>>>(z, rows, c, num_nodes, nmax, s, lds, s_full, arg);
    else k_softmax_fwd_v4<10><<<grid, 256, 0, st>>>(z, rows, c, num_nodes, nmax, s, lds, s_full, arg);
  } else if (c <= 128)
    k_softmax_fwd<4><<<grid, 256, 0, st>>>(z, rows, c, num_nodes, nmax, s, lds, s_full, arg);
  else if (c <= 512)
    k_softmax_fwd<16><<<grid, 256, 0, st>>>(z, rows, c, num_nodes, nmax, s, lds, s_full, arg);
  else
    k_softmax_fwd<40><<<grid, 256, 0, st>>>(z, rows, c, num_nodes, nmax, s, lds, s_full, arg);
  CGC_CHECK_LAUNCH("k_softmax_fwd");
  return 0;
}

extern "C" int cgc_softmax_bwd(const float* gs, const float* s, int64_t lds, int64_t rows, int32_t c, float* gz,
                               int64_t ldgz, void* stream) {
  CGC_REQUIRE(gs && s && gz && rows > 0 && c > 0 && lds >= c && ldgz >= c, CGC_ERR_ARG, "cgc_softmax_bwd: bad args");
  CGC_REQUIRE(c <= 1280, CGC_ERR_UNSUPPORTED, "cgc_softmax_bwd: c=%d > 1280", c);
  const unsigned grid = (unsigned)ceil_div(rows, 8);
  cudaStream_t st = as_stream(stream);
  if (c <= 128) k_softmax_bwd<4><<<grid, 256, 0, st>>>(gs, s, lds, rows, c, gz, ldgz);
  else if (c <= 512) k_softmax_bwd<16><<<grid, 256, 0, st>>>(gs, s, lds, rows, c, gz, ldgz);
  else k_softmax_bwd<40><<<grid, 256, 0, st>>>(gs, s, lds, rows, c, gz, ldgz);
  CGC_CHECK_LAUNCH("k_softmax_bwd");
  return 0;
}

extern "C" int cgc_row_inv_clamped_sum(const float* adj, int64_t rows, int32_t n, float* rs, float* rsum,
                                       void* stream) {
  CGC_REQUIRE(adj && rs && rsum && rows > 0 && n > 0, CGC_ERR_ARG, "cgc_row_inv_clamped_sum: bad args");
  k_row_inv_clamped_sum<<<(unsigned)ceil_div(rows, 8), 256, 0, as_stream(stream)>>>(adj, rows, n, rs, rsum);
  CGC_CHECK_LAUNCH("k_row_inv_clamped_sum");
  return 0;
}

extern "C" int cgc_add_rs_grad(float* d_adj, const float* rsum, const float* d_rs, int64_t rows, int32_t n,
                               void* stream) {
  CGC_REQUIRE(d_adj && rsum && d_rs && rows > 0 && n > 0, CGC_ERR_ARG, "cgc_add_rs_grad: bad args");
  k_add_rs_grad<<<(unsigned)ceil_div(rows, 8), 256, 0, as_stream(stream)>>>(d_adj, rsum, d_rs, rows, n);
  CGC_CHECK_LAUNCH("k_add_rs_grad");
  return 0;
}

extern "C" int cgc_dense_renorm_fwd(const float* adj, int32_t num_graphs, int32_t n, double p,
                                    const int32_t* num_nodes, float* out, float* rho, void* stream) {
  CGC_REQUIRE(adj && out && rho && num_graphs > 0 && n > 0, CGC_ERR_ARG, "cgc_dense_renorm_fwd: bad args");
  int64_t rows = (int64_t)num_graphs * n;
  k_dense_renorm_fwd<<<(unsigned)ceil_div(rows, 8), 256, 0, as_stream(stream)>>>(adj, rows, n, (float)p,
                                                                                 (float)(1.0 - p), num_nodes, out, rho);
  CGC_CHECK_LAUNCH("k_dense_renorm_fwd");
  return 0;
}

extern "C" int cgc_dense_renorm_bwd(const float* g, const float* out, const float* rho, int32_t num_graphs,
                                    int32_t n, double p, const int32_t* num_nodes, float* g_adj, void* stream) {
  CGC_REQUIRE(g && out && rho && g_adj && num_graphs > 0 && n > 0, CGC_ERR_ARG, "cgc_dense_renorm_bwd: bad args");
  int64_t rows = (int64_t)num_graphs * n;
  k_dense_renorm_bwd<<<(unsigned)ceil_div(rows, 8), 256, 0, as_stream(stream)>>>(g, out, rho, rows, n,
                                                                                 (float)(1.0 - p), num_nodes, g_adj);
  CGC_CHECK_LAUNCH("k_dense_renorm_bwd");
  return 0;
}
