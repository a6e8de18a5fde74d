// Column-wise operators: BatchNorm1d over all padded rows (reference GNN_Module.bn,
// model/network.py:101-107), bias gradients (column sums) and the max read-out (network.py:264).
// Column reductions are two-phase and deterministic: per-CTA partials in the workspace, then a
// fixed-order finalize in double precision.
#include "common.cuh"

namespace cgc {

constexpr int kRedX = 32, kRedY = 8;
enum { RED_STATS = 0, RED_BN_BWD = 1, RED_COLSUM = 2 };

// partial[(rb * 2 + which) * c + col]
template <int MODE>
__global__ void __launch_bounds__(kRedX* kRedY)
k_colreduce(const float* __restrict__ a, int64_t lda, const float* __restrict__ y, int64_t ldy,
            const float* __restrict__ mean, const float* __restrict__ invstd,
            const int32_t* __restrict__ num_nodes, int nmax, int64_t rows, int c, int64_t rows_per_block,
            float* __restrict__ partial) {
  __shared__ float s1[kRedY][kRedX], s2[kRedY][kRedX];
  const int col = blockIdx.x * kRedX + threadIdx.x;
  const int64_t rb = blockIdx.y;
  const int64_t r_begin = rb * rows_per_block;
  const int64_t r_end = min(rows, r_begin + rows_per_block);
  float acc1 = 0.f, acc2 = 0.f;
  if (col < c) {
    float mu = 0.f, is = 1.f;
    if (MODE == RED_BN_BWD) { mu = mean[col]; is = invstd[col]; }
    // kUnroll independent loads in flight per thread (the loop is latency-bound otherwise)
    constexpr int kUnroll = 8;
    for (int64_t r0 = r_begin + threadIdx.y; r0 < r_end; r0 += (int64_t)kRedY * kUnroll) {
      float va[kUnroll], vy[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int64_t r = r0 + (int64_t)u * kRedY;
        const bool ok = r < r_end && (MODE == RED_STATS || row_valid(num_nodes, nmax, r));
        va[u] = ok ? __ldg(a + r * lda + col) : 0.f;
        vy[u] = (MODE == RED_BN_BWD && ok) ? __ldg(y + r * ldy + col) : mu;
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        acc1 += va[u];
        if (MODE == RED_STATS) acc2 = fmaf(va[u], va[u], acc2);
        if (MODE == RED_BN_BWD) acc2 = fmaf(va[u], (vy[u] - mu) * is, acc2);
      }
    }
  }
  s1[threadIdx.y][threadIdx.x] = acc1;
  s2[threadIdx.y][threadIdx.x] = acc2;
  __syncthreads();
  if (threadIdx.y == 0 && col < c) {
    float t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int i = 0; i < kRedY; ++i) { t1 += s1[i][threadIdx.x]; t2 += s2[i][threadIdx.x]; }
    partial[(rb * 2 + 0) * c + col] = t1;
    partial[(rb * 2 + 1) * c + col] = t2;
  }
}

// float4 twin of k_colreduce for wide, 16-byte aligned matrices (c % 4 == 0): one thread owns four columns, so a row
// costs ONE validity test (a 32-bit division) per 16 bytes instead of one per 4 -- the scalar kernel spent 57 % of
// the SM's issue slots on that arithmetic and streamed at 0.38 - 0.46 of the HBM peak.  Same partial layout, same
// fixed-order finalize.
constexpr int kRedUnroll4 = 4;
template <int MODE>
__global__ void __launch_bounds__(kRedX* kRedY)
k_colreduce_v4(const float* __restrict__ a, int64_t lda, const float* __restrict__ y, int64_t ldy,
               const float* __restrict__ mean, const float* __restrict__ invstd,
               const int32_t* __restrict__ num_nodes, int nmax, int64_t rows, int c, int64_t rows_per_block,
               float* __restrict__ partial) {
  __shared__ float4 s1[kRedY][kRedX], s2[kRedY][kRedX];
  const int ch = blockIdx.x * kRedX + threadIdx.x;
  const int c4 = c >> 2;
  const int64_t rb = blockIdx.y;
  const int64_t r_begin = rb * rows_per_block;
  const int64_t r_end = min(rows, r_begin + rows_per_block);
  float4 acc1 = make_float4(0.f, 0.f, 0.f, 0.f), acc2 = acc1;
  if (ch < c4) {
    float4 mu = acc1, is = make_float4(1.f, 1.f, 1.f, 1.f);
    if (MODE == RED_BN_BWD) {
      mu = __ldg(reinterpret_cast<const float4*>(mean) + ch);
      is = __ldg(reinterpret_cast<const float4*>(invstd) + ch);
    }
    for (int64_t r0 = r_begin + threadIdx.y; r0 < r_end; r0 += (int64_t)kRedY * kRedUnroll4) {
      float4 va[kRedUnroll4], vy[kRedUnroll4];
#pragma unroll
      for (int u = 0; u < kRedUnroll4; ++u) {
        const int64_t r = r0 + (int64_t)u * kRedY;
        const bool ok = r < r_end && (MODE == RED_STATS || row_valid(num_nodes, nmax, r));
        va[u] = ok ? __ldg(reinterpret_cast<const float4*>(a + r * lda) + ch) : make_float4(0.f, 0.f, 0.f, 0.f);
        vy[u] = (MODE == RED_BN_BWD && ok) ? __ldg(reinterpret_cast<const float4*>(y + r * ldy) + ch) : mu;
      }
#pragma unroll
      for (int u = 0; u < kRedUnroll4; ++u) {
        acc1.x += va[u].x; acc1.y += va[u].y; acc1.z += va[u].z; acc1.w += va[u].w;
        if (MODE == RED_STATS) {
          acc2.x = fmaf(va[u].x, va[u].x, acc2.x); acc2.y = fmaf(va[u].y, va[u].y, acc2.y);
          acc2.z = fmaf(va[u].z, va[u].z, acc2.z); acc2.w = fmaf(va[u].w, va[u].w, acc2.w);
        }
        if (MODE == RED_BN_BWD) {
          acc2.x = fmaf(va[u].x, (vy[u].x - mu.x) * is.x, acc2.x); acc2.y = fmaf(va[u].y, (vy[u].y - mu.y) * is.y, acc2.y);
          acc2.z = fmaf(va[u].z, (vy[u].z - mu.z) * is.z, acc2.z); acc2.w = fmaf(va[u].w, (vy[u].w - mu.w) * is.w, acc2.w);
        }
      }
    }
  }
  s1[threadIdx.y][threadIdx.x] = acc1;
  s2[threadIdx.y][threadIdx.x] = acc2;
  __syncthreads();
  if (threadIdx.y == 0 && ch < c4) {
    float4 t1 = make_float4(0.f, 0.f, 0.f, 0.f), t2 = t1;
#pragma unroll
    for (int i = 0; i < kRedY; ++i) {
      const float4 p = s1[i][threadIdx.x], q = s2[i][threadIdx.x];
      t1.x += p.x; t1.y += p.y; t1.z += p.z; t1.w += p.w;
      t2.x += q.x; t2.y += q.y; t2.z += q.z; t2.w += q.w;
    }
    *reinterpret_cast<float4*>(partial + (rb * 2 + 0) * c + 4 * ch) = t1;     // c % 4 == 0: 16-byte aligned
    *reinterpret_cast<float4*>(partial + (rb * 2 + 1) * c + 4 * ch) = t2;
  }
}

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// one warp per column: lanes stride over the row-block partials (fixed order -> deterministic)
template <typename P>
__device__ __forceinline__ void reduce_partials(const P* __restrict__ partial, int nrb, int c, int col,
                                                double& t1, double& t2) {
  const int lane = threadIdx.x & 31;
  t1 = 0.0;
  t2 = 0.0;
  for (int rb = lane; rb < nrb; rb += 32) {
    t1 += (double)partial[(rb * 2 + 0) * (int64_t)c + col];
    t2 += (double)partial[(rb * 2 + 1) * (int64_t)c + col];
  }
  t1 = warp_sum_f64(t1);
  t2 = warp_sum_f64(t2);
}

template <typename P>
__global__ void __launch_bounds__(256)
k_bn_finalize(const P* __restrict__ partial, int nrb, int c, int64_t rows, float* __restrict__ running_mean,
              float* __restrict__ running_var, float momentum, float eps, float* __restrict__ mean,
              float* __restrict__ invstd) {
  int col = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (col >= c) return;
  double t1, t2;
  reduce_partials(partial, nrb, c, col, t1, t2);
  if ((threadIdx.x & 31) != 0) return;
  double mu = t1 / (double)rows;
  double var = t2 / (double)rows - mu * mu;
  if (var < 0.0) var = 0.0;
  mean[col] = (float)mu;
  invstd[col] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) running_mean[col] = (1.0f - momentum) * running_mean[col] + momentum * (float)mu;
  if (running_var) {
    double unbiased = rows > 1 ? var * (double)rows / (double)(rows - 1) : var;
    running_var[col] = (1.0f - momentum) * running_var[col] + momentum * (float)unbiased;
  }
}

// out1[col] = sum partial[., 0, col]; out2[col] = sum partial[., 1, col]   (out2 nullable)
__global__ void __launch_bounds__(256)
k_sum_finalize(const float* __restrict__ partial, int nrb, int c, float* __restrict__ out1,
               float* __restrict__ out2) {
  int col = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (col >= c) return;
  double t1, t2;
  reduce_partials(partial, nrb, c, col, t1, t2);
  if ((threadIdx.x & 31) != 0) return;
  out1[col] = (float)t1;
  if (out2) out2[col] = (float)t2;
}

__global__ void k_bn_eval_stats(const float* __restrict__ rm, const float* __restrict__ rv, int c, float eps,
                                float* __restrict__ mean, float* __restrict__ invstd) {
  int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= c) return;
  mean[col] = rm[col];
  invstd[col] = 1.0f / sqrtf(rv[col] + eps);
}

// out = mask * ((y - mean) * invstd * gamma + beta)      (mean == nullptr: out = mask * y)
template <typename IdxT>
__global__ void __launch_bounds__(256)
k_bn_apply(const float* __restrict__ y, int64_t ldy, const float* __restrict__ mean,
           const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
           const int32_t* __restrict__ num_nodes, int nmax, int64_t rows, int c, float* __restrict__ out,
           int64_t ldo) {
  const IdxT total = (IdxT)(rows * c);
  for (IdxT idx = (IdxT)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (IdxT)gridDim.x * blockDim.x) {
    const IdxT r = idx / (IdxT)c;
    const int col = (int)(idx - r * (IdxT)c);
    float v = 0.f;
    if (row_valid(num_nodes, nmax, (int64_t)r)) {
      v = y[(int64_t)r * ldy + col];
      if (mean != nullptr) v = (v - mean[col]) * invstd[col] * gamma[col] + beta[col];
    }
    out[(int64_t)r * ldo + col] = v;
  }
}

// float4 variant for wide, 16-byte aligned matrices: block (32 chunks x 8 rows), kRowUnroll rows in flight
constexpr int kRowUnroll = 4;
__global__ void __launch_bounds__(256)
k_bn_apply_v4(const float* __restrict__ y, int64_t ldy, const float* __restrict__ mean,
              const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
              const int32_t* __restrict__ num_nodes, int nmax, int64_t rows, int c4, float* __restrict__ out,
              int64_t ldo) {
  const int ch = blockIdx.x * 32 + threadIdx.x;
  if (ch >= c4) return;
  if (mean != nullptr) {
    const float4 mu = __ldg(reinterpret_cast<const float4*>(mean) + ch);
    const float4 is = __ldg(reinterpret_cast<const float4*>(invstd) + ch);
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + ch);
    const float4 be = __ldg(reinterpret_cast<const float4*>(beta) + ch);
    // (v - mu) * is * ga + be, evaluated in that order to stay bit-identical with the scalar kernel
    for (int64_t r0 = (int64_t)blockIdx.y * 8 + threadIdx.y; r0 < rows; r0 += (int64_t)gridDim.y * 8 * kRowUnroll) {
      float4 v[kRowUnroll];
#pragma unroll
      for (int u = 0; u < kRowUnroll; ++u) {
        const int64_t r = r0 + (int64_t)u * gridDim.y * 8;
        if (r < rows) v[u] = __ldg(reinterpret_cast<const float4*>(y + r * ldy) + ch);
      }
#pragma unroll
      for (int u = 0; u < kRowUnroll; ++u) {
        const int64_t r = r0 + (int64_t)u * gridDim.y * 8;
        if (r >= rows) continue;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row_valid(num_nodes, nmax, r)) {
          o.x = (v[u].x - mu.x) * is.x * ga.x + be.x;
          o.y = (v[u].y - mu.y) * is.y * ga.y + be.y;
          o.z = (v[u].z - mu.z) * is.z * ga.z + be.z;
          o.w = (v[u].w - mu.w) * is.w * ga.w + be.w;
        }
        *(reinterpret_cast<float4*>(out + r * ldo) + ch) = o;
      }
    }
    return;
  }
  for (int64_t r0 = (int64_t)blockIdx.y * 8 + threadIdx.y; r0 < rows; r0 += (int64_t)gridDim.y * 8 * kRowUnroll) {
#pragma unroll
    for (int u = 0; u < kRowUnroll; ++u) {
      const int64_t r = r0 + (int64_t)u * gridDim.y * 8;
      if (r >= rows) continue;
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row_valid(num_nodes, nmax, r)) o = __ldg(reinterpret_cast<const float4*>(y + r * ldy) + ch);
      *(reinterpret_cast<float4*>(out + r * ldo) + ch) = o;
    }
  }
}

// gy = gamma*invstd*(g - d_beta/rows - xhat*d_gamma/rows)   (training) | gamma*invstd*g   (eval)
template <typename IdxT>
__global__ void __launch_bounds__(256)
k_bn_bwd_apply(const float* __restrict__ gx, int64_t ldgx, const float* __restrict__ y, int64_t ldy,
               const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
               const float* __restrict__ d_gamma, const float* __restrict__ d_beta,
               const int32_t* __restrict__ num_nodes, int nmax, int64_t rows, int c, int training,
               float* __restrict__ gy, int64_t ldgy) {
  const IdxT total = (IdxT)(rows * c);
  const float inv_rows = 1.0f / (float)rows;
  for (IdxT idx = (IdxT)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (IdxT)gridDim.x * blockDim.x) {
    const IdxT r = idx / (IdxT)c;
    const int col = (int)(idx - r * (IdxT)c);
    const float g = row_valid(num_nodes, nmax, (int64_t)r) ? gx[(int64_t)r * ldgx + col] : 0.f;
    const float is = invstd[col];
    float v;
    if (training) {
      const float xhat = (y[(int64_t)r * ldy + col] - mean[col]) * is;
      v = gamma[col] * is * (g - d_beta[col] * inv_rows - xhat * d_gamma[col] * inv_rows);
    } else {
      v = gamma[col] * is * g;
    }
    gy[(int64_t)r * ldgy + col] = v;
  }
}

__device__ __forceinline__ float bn_bwd_elem(float g, float yv, float mu, float is, float ga, float dg, float db,
                                             float inv_rows, int training) {
  if (training) {
    const float xhat = (yv - mu) * is;
    return ga * is * (g - db * inv_rows - xhat * dg * inv_rows);
  }
  return ga * is * g;
}

__global__ void __launch_bounds__(256)
k_bn_bwd_apply_v4(const float* __restrict__ gx, int64_t ldgx, const float* __restrict__ y, int64_t ldy,
                  const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
                  const float* __restrict__ d_gamma, const float* __restrict__ d_beta,
                  const int32_t* __restrict__ num_nodes, int nmax, int64_t rows, int c4, int training,
                  float* __restrict__ gy, int64_t ldgy) {
  const int ch = blockIdx.x * 32 + threadIdx.x;
  if (ch >= c4) return;
  const float inv_rows = 1.0f / (float)rows;
  const float4 mu = __ldg(reinterpret_cast<const float4*>(mean) + ch);
  const float4 is = __ldg(reinterpret_cast<const float4*>(invstd) + ch);
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + ch);
  const float4 dg = __ldg(reinterpret_cast<const float4*>(d_gamma) + ch);
  const float4 db = __ldg(reinterpret_cast<const float4*>(d_beta) + ch);
  for (int64_t r0 = (int64_t)blockIdx.y * 8 + threadIdx.y; r0 < rows; r0 += (int64_t)gridDim.y * 8 * kRowUnroll) {
    float4 g[kRowUnroll], yv[kRowUnroll];
#pragma unroll
    for (int u = 0; u < kRowUnroll; ++u) {
      const int64_t r = r0 + (int64_t)u * gridDim.y * 8;
      g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      yv[u] = mu;
      if (r < rows) {
        if (row_valid(num_nodes, nmax, r)) g[u] = __ldg(reinterpret_cast<const float4*>(gx + r * ldgx) + ch);
        if (training) yv[u] = __ldg(reinterpret_cast<const float4*>(y + r * ldy) + ch);
      }
    }
#pragma unroll
    for (int u = 0; u < kRowUnroll; ++u) {
      const int64_t r = r0 + (int64_t)u * gridDim.y * 8;
      if (r >= rows) continue;
      float4 o;
      o.x = bn_bwd_elem(g[u].x, yv[u].x, mu.x, is.x, ga.x, dg.x, db.x, inv_rows, training);
      o.y = bn_bwd_elem(g[u].y, yv[u].y, mu.y, is.y, ga.y, dg.y, db.y, inv_rows, training);
      o.z = bn_bwd_elem(g[u].z, yv[u].z, mu.z, is.z, ga.z, dg.z, db.z, inv_rows, training);
      o.w = bn_bwd_elem(g[u].w, yv[u].w, mu.w, is.w, ga.w, dg.w, db.w, inv_rows, training);
      *(reinterpret_cast<float4*>(gy + r * ldgy) + ch) = o;
    }
  }
}

// torch.max(x.view(B, nmax, c), dim=1): first maximal row wins ties
__global__ void __launch_bounds__(kRedX* kRedY)
k_colmax_fwd(const float* __restrict__ x, int nmax, int c, float* __restrict__ out, int32_t* __restrict__ arg) {
  __shared__ float sv[kRedY][kRedX];
  __shared__ int si[kRedY][kRedX];
  const int col = blockIdx.x * kRedX + threadIdx.x;
  const int g = blockIdx.y;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  if (col < c) {
    const float* base = x + (int64_t)g * nmax * c + col;
    for (int r = threadIdx.y; r < nmax; r += kRedY) {
      float v = __ldg(base + (int64_t)r * c);
      if (v > best) { best = v; bi = r; }
    }
  }
  sv[threadIdx.y][threadIdx.x] = best;
  si[threadIdx.y][threadIdx.x] = bi;
  __syncthreads();
  if (threadIdx.y == 0 && col < c) {
#pragma unroll
    for (int i = 1; i < kRedY; ++i) {
      float v = sv[i][threadIdx.x];
      int ii = si[i][threadIdx.x];
      if (v > best || (v == best && ii < bi)) { best = v; bi = ii; }
    }
    out[(int64_t)g * c + col] = best;
    arg[(int64_t)g * c + col] = bi == 0x7fffffff ? 0 : bi;
  }
}

__global__ void k_colmax_bwd(const float* __restrict__ g, const int32_t* __restrict__ arg, int num_graphs,
                             int nmax, int c, float* __restrict__ gx) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)num_graphs * c) return;
  int64_t gi = idx / c;
  int col = (int)(idx - gi * c);
  gx[(gi * nmax + arg[idx]) * c + col] += g[idx];
}

// ---------------------------------------------------------------------------------------------
// ------------------------------------------------------------------ fused narrow SAGE layer, forward
// y = act(l2normalize(sa W + b) * mask) for cout <= 32, cin <= 64, PLUS the BatchNorm column sums of y in the same
// pass: one thread per row (the warp-per-row epilogue spent its time in shuffles: 36 us for a 9 MB tensor), the
// 256-row tile staged through shared memory so that both the sa reads and the y writes are coalesced, W read as
// warp-wide broadcasts.  Each CTA leaves sum / sum-of-squares partials in the layout k_bn_finalize reads, so the
// separate statistics pass over y (k_colreduce<stats>) disappears.
constexpr int kNfThreads = 256;

__global__ void __launch_bounds__(kNfThreads)
k_sage_fwd_narrow(const float* __restrict__ sa, int64_t ldsa, const float* __restrict__ w, const float* __restrict__ bias,
                  const int32_t* __restrict__ num_nodes, int nmax, int64_t rows, int cin, int cout, int act,
                  float* __restrict__ y, int64_t ldy, double* __restrict__ partial) {
  extern __shared__ float nf_smem[];
  const int sa_stride = cin | 1;                         // odd strides: conflict-free row-per-thread access
  const int sy_stride = cout | 1;
  float* s_w = nf_smem;                                  // [cin][32]
  float* s_b = s_w + cin * 32;                           // [32]
  float* s_a = s_b + 32;                                 // [256][sa_stride]
  float* s_y = s_a + kNfThreads * sa_stride;             // [256][sy_stride]
  __shared__ double s_p[(kNfThreads / 32) * 2 * 32];     // [8 warps][2][32] column sums (double: var = E[y^2] - mean^2 cancels)
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  for (int i = t; i < cin * 32; i += kNfThreads) {
    const int k = i >> 5, c = i & 31;
    s_w[i] = c < cout ? __ldg(w + k * cout + c) : 0.f;
  }
  if (t < 32) s_b[t] = (bias && t < cout) ? __ldg(bias + t) : 0.f;
  double tot1 = 0.0, tot2 = 0.0;                         // this thread's column (lane < cout) over warp's rows, all tiles
  const int64_t ntiles = (rows + kNfThreads - 1) / kNfThreads;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t r0 = tile * kNfThreads;
    const int nr = (int)min((int64_t)kNfThreads, rows - r0);
    __syncthreads();                                     // previous tile's s_a / s_y consumed (and s_w staged)
    for (int i = t; i < nr * cin; i += kNfThreads) {     // coalesced when ldsa == cin (row-major contiguous tile)
      const int rr = i / cin, k = i - rr * cin;
      s_a[rr * sa_stride + k] = __ldg(sa + (r0 + rr) * ldsa + k);
    }
    __syncthreads();
    if (t < nr) {
      float h[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) h[c] = s_b[c];
      const float* arow = s_a + t * sa_stride;
      for (int k = 0; k < cin; ++k) {
        const float a = arow[k];
        const float4* wr = reinterpret_cast<const float4*>(s_w + k * 32);
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 wv = wr[c4];
          h[4 * c4] = fmaf(a, wv.x, h[4 * c4]); h[4 * c4 + 1] = fmaf(a, wv.y, h[4 * c4 + 1]);
          h[4 * c4 + 2] = fmaf(a, wv.z, h[4 * c4 + 2]); h[4 * c4 + 3] = fmaf(a, wv.w, h[4 * c4 + 3]);
        }
      }
      float n2 = 0.f;
#pragma unroll
      for (int c = 0; c < 32; ++c) n2 = fmaf(h[c], h[c], n2);          // padded columns are exactly 0
      const float inv = 1.0f / fmaxf(sqrtf(n2), kL2NormEps);
      const bool valid = row_valid(num_nodes, nmax, r0 + t);
      float* yrow = s_y + t * sy_stride;
#pragma unroll
      for (int c = 0; c < 32; ++c)
        if (c < cout) yrow[c] = valid ? act_fwd(act, h[c] * inv) : 0.f;
    }
    __syncthreads();
    for (int i = t; i < nr * cout; i += kNfThreads) {    // coalesced write-out
      const int rr = i / cout, c = i - rr * cout;
      y[(r0 + rr) * ldy + c] = s_y[rr * sy_stride + c];
    }
    if (lane < cout) {                                   // column sums over this warp's 32 rows of the tile
      const int rbeg = warp * 32, rend = min(nr, rbeg + 32);
      for (int rr = rbeg; rr < rend; ++rr) {
        const double v = (double)s_y[rr * sy_stride + lane];
        tot1 += v;
        tot2 = fma(v, v, tot2);
      }
    }
  }
  s_p[(warp * 2 + 0) * 32 + lane] = tot1;
  s_p[(warp * 2 + 1) * 32 + lane] = tot2;
  __syncthreads();
  if (t < 2 * cout) {
    const int which = t / cout, c = t - which * cout;
    double v = 0.0;
#pragma unroll
    for (int wv = 0; wv < kNfThreads / 32; ++wv) v += s_p[(wv * 2 + which) * 32 + c];
    partial[((int64_t)blockIdx.x * 2 + which) * cout + c] = v;
  }
}

// BatchNorm-backward column sums for narrow matrices (c <= 32): sum_r g and sum_r g * xhat with g = gx * mask.
// The generic k_colreduce gives 32 lanes to the columns of ONE row (20 of them busy, 80-byte requests); here a
// CTA stages a 256-row tile of gx and y with coalesced loads and each warp reduces its 32 rows per column.
__global__ void __launch_bounds__(kNfThreads)
k_bn_bwd_sums_narrow(const float* __restrict__ gx, int64_t ldgx, const float* __restrict__ y, int64_t ldy,
                     const float* __restrict__ mean, const float* __restrict__ invstd,
                     const int32_t* __restrict__ num_nodes, int nmax, int64_t rows, int c,
                     float* __restrict__ partial) {
  constexpr int kRows = 128;                       // rows per tile (2 x 128 x 33 floats of static shared memory)
  __shared__ float s_g[kRows * 33], s_x[kRows * 33];
  __shared__ float s_p[(kNfThreads / 32) * 2 * 32];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const float mu = lane < c ? __ldg(mean + lane) : 0.f, is = lane < c ? __ldg(invstd + lane) : 0.f;
  float tot1 = 0.f, tot2 = 0.f;
  const int64_t ntiles = (rows + kRows - 1) / kRows;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t r0 = tile * kRows;
    const int nr = (int)min((int64_t)kRows, rows - r0);
    __syncthreads();
    for (int i = t; i < nr * c; i += kNfThreads) {
      const int rr = i / c, k = i - rr * c;
      const int64_t r = r0 + rr;
      s_g[rr * 33 + k] = __ldg(gx + r * ldgx + k);
      s_x[rr * 33 + k] = __ldg(y + r * ldy + k);
    }
    __syncthreads();
    if (lane < c) {
      const int rbeg = warp * (kRows / 8), rend = min(nr, rbeg + kRows / 8);
      for (int rr = rbeg; rr < rend; ++rr) {
        const float g = row_valid(num_nodes, nmax, r0 + rr) ? s_g[rr * 33 + lane] : 0.f;
        tot1 += g;
        tot2 = fmaf(g, (s_x[rr * 33 + lane] - mu) * is, tot2);
      }
    }
  }
  s_p[(warp * 2 + 0) * 32 + lane] = tot1;
  s_p[(warp * 2 + 1) * 32 + lane] = tot2;
  __syncthreads();
  if (t < 2 * c) {
    const int which = t / c, col = t - which * c;
    float v = 0.f;
#pragma unroll
    for (int wv = 0; wv < kNfThreads / 32; ++wv) v += s_p[(wv * 2 + which) * 32 + col];
    partial[((int64_t)blockIdx.x * 2 + which) * c + col] = v;
  }
}

// vec: the float4 kernel (one thread = four columns)
static void reduce_geometry(int64_t rows, int c, int& nrb, int64_t& rows_per_block, dim3& grid, bool vec = false) {
  int colblocks = (int)ceil_div(vec ? c / 4 : c, kRedX);
  int64_t want = ceil_div((int64_t)kNumSMs * (vec ? 8 : 6), colblocks);
  int64_t max_by_rows = ceil_div(rows, kRedY * (vec ? kRedUnroll4 : 8));      // at least one unrolled batch per thread
  if (want > max_by_rows) want = max_by_rows;
  if (want < 1) want = 1;
  if (want > 512) want = 512;
  rows_per_block = ceil_div(rows, want);
  nrb = (int)ceil_div(rows, rows_per_block);
  grid = dim3((unsigned)colblocks, (unsigned)nrb);
}

}  // namespace cgc

namespace cgc {
// shared with sage_layer.cu: finalize of per-block [2][c] double partial sums
int bn_finalize_partials(const double* partial, int blocks, int c, int64_t rows, float* running_mean,
                         float* running_var, float momentum, float eps, float* mean, float* invstd, cudaStream_t st) {
  k_bn_finalize<double><<<(unsigned)ceil_div(c, 8), 256, 0, st>>>(partial, blocks, c, rows, running_mean, running_var,
                                                                 momentum, eps, mean, invstd);
  CGC_CHECK_LAUNCH("k_bn_finalize");
  return 0;
}
}  // namespace cgc

using namespace cgc;

// wide aligned matrices take the float4 kernel
static bool reduce_vec_ok(int c, const float* a, int64_t lda, const float* y = nullptr, int64_t ldy = 0,
                          const float* mean = nullptr, const float* invstd = nullptr) {
  return c >= 128 && c % 4 == 0 && lda % 4 == 0 && aligned16(a) && (y == nullptr || (ldy % 4 == 0 && aligned16(y))) &&
         (mean == nullptr || (aligned16(mean) && aligned16(invstd)));
}

extern "C" int cgc_bn_stats(const float* y, int64_t ldy, int64_t rows, int32_t c, float* running_mean,
                            float* running_var, float momentum, float eps, float* mean, float* invstd,
                            void* work, int64_t work_bytes, void* stream) {
  CGC_REQUIRE(y && mean && invstd && work && rows > 0 && c > 0 && ldy >= c, CGC_ERR_ARG, "cgc_bn_stats: bad args");
  int nrb;
  int64_t rpb;
  dim3 grid;
  const bool vec = reduce_vec_ok(c, y, ldy) && aligned16(work);
  reduce_geometry(rows, c, nrb, rpb, grid, vec);
  CGC_REQUIRE((int64_t)nrb * 2 * c * 4 <= work_bytes, CGC_ERR_WORKSPACE, "cgc_bn_stats: workspace too small");
  cudaStream_t st = as_stream(stream);
  float* partial = reinterpret_cast<float*>(work);
  if (vec)
    k_colreduce_v4<RED_STATS><<<grid, dim3(kRedX, kRedY), 0, st>>>(y, ldy, nullptr, 0, nullptr, nullptr, nullptr, 0,
                                                                   rows, c, rpb, partial);
  else
    k_colreduce<RED_STATS><<<grid, dim3(kRedX, kRedY), 0, st>>>(y, ldy, nullptr, 0, nullptr, nullptr, nullptr, 0, rows,
                                                                c, rpb, partial);
  CGC_CHECK_LAUNCH("k_colreduce<stats>");
  k_bn_finalize<float><<<(unsigned)ceil_div(c, 8), 256, 0, st>>>(partial, nrb, c, rows, running_mean, running_var,
                                                            momentum, eps, mean, invstd);
  CGC_CHECK_LAUNCH("k_bn_finalize");
  return 0;
}

extern "C" int cgc_bn_eval_stats(const float* running_mean, const float* running_var, int32_t c, float eps,
                                 float* mean, float* invstd, void* stream) {
  CGC_REQUIRE(running_mean && running_var && mean && invstd && c > 0, CGC_ERR_ARG, "cgc_bn_eval_stats: bad args");
  k_bn_eval_stats<<<(unsigned)ceil_div(c, 128), 128, 0, as_stream(stream)>>>(running_mean, running_var, c, eps,
                                                                             mean, invstd);
  CGC_CHECK_LAUNCH("k_bn_eval_stats");
  return 0;
}

// grid for the float4 row-streaming kernels: x = 32-chunk column groups, y = row blocks (~8 CTAs per SM)
static dim3 wide_grid(int64_t rows, int c4) {
  int gx = (int)ceil_div(c4, 32);
  int64_t gy = ceil_div((int64_t)kNumSMs * 8, gx);
  int64_t maxy = ceil_div(rows, 8 * kRowUnroll);
  if (gy > maxy) gy = maxy;
  if (gy < 1) gy = 1;
  if (gy > 65535) gy = 65535;
  return dim3((unsigned)gx, (unsigned)gy);
}

static unsigned elementwise_grid(int64_t total) {
  int64_t blocks = ceil_div(total, 256);
  int64_t cap = (int64_t)kNumSMs * 32;
  return (unsigned)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

extern "C" int cgc_bn_apply(const float* y, int64_t ldy, const float* mean, const float* invstd,
                            const float* gamma, const float* beta, const int32_t* num_nodes, int32_t nmax,
                            int64_t rows, int32_t c, float* out, int64_t ldo, void* stream) {
  CGC_REQUIRE(y && out && rows > 0 && c > 0 && ldy >= c && ldo >= c, CGC_ERR_ARG, "cgc_bn_apply: bad args");
  CGC_REQUIRE(mean == nullptr || (invstd && gamma && beta), CGC_ERR_ARG, "cgc_bn_apply: incomplete BN params");
  CGC_REQUIRE(num_nodes == nullptr || nmax > 0, CGC_ERR_ARG, "cgc_bn_apply: mask without nmax");
  cudaStream_t st = as_stream(stream);
  const bool vec = c >= 64 && c % 4 == 0 && ldy % 4 == 0 && ldo % 4 == 0 && aligned16(y) && aligned16(out) &&
                   (mean == nullptr || (aligned16(mean) && aligned16(invstd) && aligned16(gamma) && aligned16(beta)));
  if (vec) {
    k_bn_apply_v4<<<wide_grid(rows, c / 4), dim3(32, 8), 0, st>>>(y, ldy, mean, invstd, gamma, beta, num_nodes, nmax,
                                                                 rows, c / 4, out, ldo);
  } else if (rows * c < (1ll << 31)) {
    k_bn_apply<uint32_t><<<elementwise_grid(rows * c), 256, 0, st>>>(y, ldy, mean, invstd, gamma, beta, num_nodes,
                                                                     nmax, rows, c, out, ldo);
  } else {
    k_bn_apply<int64_t><<<elementwise_grid(rows * c), 256, 0, st>>>(y, ldy, mean, invstd, gamma, beta, num_nodes,
                                                                    nmax, rows, c, out, ldo);
  }
  CGC_CHECK_LAUNCH("k_bn_apply");
  return 0;
}

extern "C" int cgc_bn_bwd(const float* gx, int64_t ldgx, const float* y, int64_t ldy, const float* mean,
                          const float* invstd, const float* gamma, const int32_t* num_nodes, int32_t nmax,
                          int64_t rows, int32_t c, int training, float* gy, int64_t ldgy, float* d_gamma,
                          float* d_beta, void* work, int64_t work_bytes, void* stream) {
  CGC_REQUIRE(gx && y && mean && invstd && gamma && d_gamma && d_beta && work && rows > 0 && c > 0,
              CGC_ERR_ARG, "cgc_bn_bwd: bad args");
  CGC_REQUIRE(num_nodes == nullptr || nmax > 0, CGC_ERR_ARG, "cgc_bn_bwd: mask without nmax");
  int nrb;
  int64_t rpb;
  dim3 grid;
  const bool vec_red = reduce_vec_ok(c, gx, ldgx, y, ldy, mean, invstd) && aligned16(work);
  reduce_geometry(rows, c, nrb, rpb, grid, vec_red);
  cudaStream_t st = as_stream(stream);
  float* partial = reinterpret_cast<float*>(work);
  if (c <= 32) {
    int64_t blocks = ceil_div(rows, 128);
    if (blocks > (int64_t)kNumSMs * 4) blocks = (int64_t)kNumSMs * 4;
    nrb = (int)blocks;
    CGC_REQUIRE((int64_t)nrb * 2 * c * 4 <= work_bytes, CGC_ERR_WORKSPACE, "cgc_bn_bwd: workspace too small");
    k_bn_bwd_sums_narrow<<<(unsigned)blocks, kNfThreads, 0, st>>>(gx, ldgx, y, ldy, mean, invstd, num_nodes, nmax, rows,
                                                                 c, partial);
    CGC_CHECK_LAUNCH("k_bn_bwd_sums_narrow");
  } else {
    CGC_REQUIRE((int64_t)nrb * 2 * c * 4 <= work_bytes, CGC_ERR_WORKSPACE, "cgc_bn_bwd: workspace too small");
    if (vec_red)
      k_colreduce_v4<RED_BN_BWD><<<grid, dim3(kRedX, kRedY), 0, st>>>(gx, ldgx, y, ldy, mean, invstd, num_nodes, nmax,
                                                                      rows, c, rpb, partial);
    else
      k_colreduce<RED_BN_BWD><<<grid, dim3(kRedX, kRedY), 0, st>>>(gx, ldgx, y, ldy, mean, invstd, num_nodes, nmax,
                                                                   rows, c, rpb, partial);
    CGC_CHECK_LAUNCH("k_colreduce<bn_bwd>");
  }
  k_sum_finalize<<<(unsigned)ceil_div(c, 8), 256, 0, st>>>(partial, nrb, c, d_beta, d_gamma);
  CGC_CHECK_LAUNCH("k_sum_finalize");
  if (gy == nullptr) return 0;                       // sums only: the consumer folds the apply step (cgc_sage_dh_bn)
  const bool vec = c >= 64 && c % 4 == 0 && ldgx % 4 == 0 && ldy % 4 == 0 && ldgy % 4 == 0 && aligned16(gx) &&
                   aligned16(y) && aligned16(gy) && aligned16(mean) && aligned16(invstd) && aligned16(gamma) &&
                   aligned16(d_gamma) && aligned16(d_beta);
  if (vec) {
    k_bn_bwd_apply_v4<<<wide_grid(rows, c / 4), dim3(32, 8), 0, st>>>(gx, ldgx, y, ldy, mean, invstd, gamma, d_gamma,
                                                                     d_beta, num_nodes, nmax, rows, c / 4, training,
                                                                     gy, ldgy);
  } else if (rows * c < (1ll << 31)) {
    k_bn_bwd_apply<uint32_t><<<elementwise_grid(rows * c), 256, 0, st>>>(gx, ldgx, y, ldy, mean, invstd, gamma,
                                                                         d_gamma, d_beta, num_nodes, nmax, rows, c,
                                                                         training, gy, ldgy);
  } else {
    k_bn_bwd_apply<int64_t><<<elementwise_grid(rows * c), 256, 0, st>>>(gx, ldgx, y, ldy, mean, invstd, gamma,
                                                                        d_gamma, d_beta, num_nodes, nmax, rows, c,
                                                                        training, gy, ldgy);
  }
  CGC_CHECK_LAUNCH("k_bn_bwd_apply");
  return 0;
}

extern "C" int cgc_colsum(const float* x, int64_t ldx, int64_t rows, int32_t c, const int32_t* num_nodes,
                          int32_t nmax, float* out, void* work, int64_t work_bytes, void* stream) {
  CGC_REQUIRE(x && out && work && rows > 0 && c > 0 && ldx >= c, CGC_ERR_ARG, "cgc_colsum: bad args");
  CGC_REQUIRE(num_nodes == nullptr || nmax > 0, CGC_ERR_ARG, "cgc_colsum: mask without nmax");
  int nrb;
  int64_t rpb;
  dim3 grid;
  const bool vec = reduce_vec_ok(c, x, ldx) && aligned16(work);
  reduce_geometry(rows, c, nrb, rpb, grid, vec);
  CGC_REQUIRE((int64_t)nrb * 2 * c * 4 <= work_bytes, CGC_ERR_WORKSPACE, "cgc_colsum: workspace too small");
  cudaStream_t st = as_stream(stream);
  float* partial = reinterpret_cast<float*>(work);
  if (vec)
    k_colreduce_v4<RED_COLSUM><<<grid, dim3(kRedX, kRedY), 0, st>>>(x, ldx, nullptr, 0, nullptr, nullptr, num_nodes,
                                                                    nmax, rows, c, rpb, partial);
  else
    k_colreduce<RED_COLSUM><<<grid, dim3(kRedX, kRedY), 0, st>>>(x, ldx, nullptr, 0, nullptr, nullptr, num_nodes, nmax,
                                                                 rows, c, rpb, partial);
  CGC_CHECK_LAUNCH("k_colreduce<colsum>");
  k_sum_finalize<<<(unsigned)ceil_div(c, 8), 256, 0, st>>>(partial, nrb, c, out, nullptr);
  CGC_CHECK_LAUNCH("k_sum_finalize");
  return 0;
}

extern "C" int cgc_colmax_fwd(const float* x, int32_t num_graphs, int32_t nmax, int32_t c, float* out,
                              int32_t* arg, void* stream) {
  CGC_REQUIRE(x && out && arg && num_graphs > 0 && nmax > 0 && c > 0, CGC_ERR_ARG, "cgc_colmax_fwd: bad args");
  CGC_REQUIRE(num_graphs <= 65535, CGC_ERR_UNSUPPORTED, "cgc_colmax_fwd: more than 65535 graphs");
  dim3 grid((unsigned)ceil_div(c, kRedX), (unsigned)num_graphs);
  k_colmax_fwd<<<grid, dim3(kRedX, kRedY), 0, as_stream(stream)>>>(x, nmax, c, out, arg);
  CGC_CHECK_LAUNCH("k_colmax_fwd");
  return 0;
}

extern "C" int cgc_colmax_bwd(const float* g, const int32_t* arg, int32_t num_graphs, int32_t nmax, int32_t c,
                              float* gx, void* stream) {
  CGC_REQUIRE(g && arg && gx && num_graphs > 0 && nmax > 0 && c > 0, CGC_ERR_ARG, "cgc_colmax_bwd: bad args");
  k_colmax_bwd<<<(unsigned)ceil_div((int64_t)num_graphs * c, 256), 256, 0, as_stream(stream)>>>(g, arg, num_graphs,
                                                                                               nmax, c, gx);
  CGC_CHECK_LAUNCH("k_colmax_bwd");
  return 0;
}

extern "C" int cgc_sage_layer_fwd_narrow(const float* sa, int64_t ldsa, const float* w, const float* b,
                                         const int32_t* num_nodes, int32_t nmax, int64_t rows, int32_t cin,
                                         int32_t cout, int act, float* y, int64_t ldy, float* running_mean,
                                         float* running_var, float momentum, float eps, float* mean, float* invstd,
                                         void* work, int64_t work_bytes, void* stream) {
  CGC_REQUIRE(sa && w && y && mean && invstd && work && rows > 0 && cin > 0 && cout > 0 && ldsa >= cin && ldy >= cout,
              CGC_ERR_ARG, "cgc_sage_layer_fwd_narrow: bad args");
  CGC_REQUIRE(cout <= 32 && cin <= 64, CGC_ERR_UNSUPPORTED,
              "cgc_sage_layer_fwd_narrow: needs cout <= 32 and cin <= 64 (got %d, %d)", cout, cin);
  CGC_REQUIRE(num_nodes == nullptr || nmax > 0, CGC_ERR_ARG, "cgc_sage_layer_fwd_narrow: mask without nmax");
  CGC_REQUIRE(act >= CGC_ACT_RELU && act <= CGC_ACT_LEAKY_RELU, CGC_ERR_ARG,
              "cgc_sage_layer_fwd_narrow: unknown activation %d", act);
  const size_t smem = sizeof(float) * ((size_t)cin * 32 + 32 + (size_t)kNfThreads * ((cin | 1) + (cout | 1)));
  int64_t blocks = ceil_div(rows, kNfThreads);
  const int per_sm = (int)((200 * 1024) / (smem + 1024)) < 1 ? 1 : (int)((200 * 1024) / (smem + 1024));
  if (blocks > (int64_t)kNumSMs * per_sm) blocks = (int64_t)kNumSMs * per_sm;
  CGC_REQUIRE(blocks * 2 * cout * 8 <= work_bytes, CGC_ERR_WORKSPACE, "cgc_sage_layer_fwd_narrow: workspace too small");
  cudaStream_t st = as_stream(stream);
  double* partial = reinterpret_cast<double*>(work);
  CGC_CUDA(cudaFuncSetAttribute(k_sage_fwd_narrow, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_sage_fwd_narrow<<<(unsigned)blocks, kNfThreads, smem, st>>>(sa, ldsa, w, b, num_nodes, nmax, rows, cin, cout, act, y,
                                                               ldy, partial);
  CGC_CHECK_LAUNCH("k_sage_fwd_narrow");
  k_bn_finalize<double><<<(unsigned)ceil_div(cout, 8), 256, 0, st>>>(partial, (int)blocks, cout, rows, running_mean, running_var,
                                                               momentum, eps, mean, invstd);
  CGC_CHECK_LAUNCH("k_bn_finalize");
  return 0;
}
