// One narrow SAGE layer in one kernel: neighbour gather (ELL16 table, TMA-staged feature tile) -> W -> bias ->
// L2-normalise -> mask -> activation -> BatchNorm column partial sums.
//
// Reference arithmetic (model/network.py:114-116 with PyG 1.2.1 DenseSAGEConv.forward inlined):
//   out = matmul(adj, x) / adj.sum(-1).clamp(min=1);  out = matmul(out, weight) + bias;
//   out = F.normalize(out, p=2, dim=-1) * mask;  x_k = bn_k(act(out))      (this kernel stops at the batch sums)
//
// Layout of the work: CTA (g, p) owns a contiguous range of graph g's table rows and stages the graph's whole
// [nmax, CIN] input tile in shared memory (1-D bulk copies, or 2-D TMA boxes when x is a column slice of the
// concat buffer).  LPR = CIN / 4 lanes serve one row: each lane gathers one float4 of the row's neighbour sum
// exactly like k_sage_gather_ell.  Then the LPR lanes exchange their float4s (warp shuffles), so every lane holds
// the whole aggregated row, and each lane produces the output columns it owns with W kept in REGISTERS for the
// whole kernel (CIN x 4 or 5 values per lane; W never goes through shared memory, which is the saturated unit):
//   CIN = 16 (4 lanes): lane sl owns columns 4sl..4sl+3 and 16+sl;   CIN = 20 (5 lanes): columns 4sl..4sl+3.
// The squared norm is reduced over the LPR lanes, the row is scaled, masked, activated and stored, and the lane
// keeps running sums of y and y^2 of its columns; at the end they are reduced over the CTA (shuffles, then one
// shared-memory pass in double precision) into `partial[cta][2][20]` for the fixed-order finalize kernel.
#include <cuda.h>
#include <string.h>

#include "common.cuh"
#include "ell_gather.cuh"

namespace cgc {

constexpr int kLayerCout = 20;
constexpr int kLayerThreads = 384;

struct SageLayerArgs {
  const uint16_t* ell;
  const float* ellw;
  const float* ell_scale;
  const uint32_t* perm;
  const int* rowptr;
  const int* col;
  const float* val;
  const float* x;
  const float* w;            // [CIN, 20] row-major
  const float* b;            // [20] or null
  const int32_t* num_nodes;  // row mask, or null
  float* sa;                 // [rows, CIN]
  float* y;                  // [rows, ldy]
  int64_t ldy;
  double* partial;           // [grid][2][20] column sums of y and y^2, or null
  unsigned int* counter;     // completion counter of the statistics (zero on entry, reset to zero by the last CTA)
  float* running_mean;       // BatchNorm outputs written by the last CTA
  float* running_var;
  float* mean;
  float* invstd;
  float momentum, eps;
  int nmax, parts, rpp, box_rows, early, act;
  int tile_smem;             // bytes of the feature tile buffer (the table slice follows it)
};

template <int CIN, int W, bool WEIGHTED, bool STRIDED>
__global__ void __launch_bounds__(kLayerThreads, 1)
k_sage_layer_ell(const __grid_constant__ CUtensorMap map_x, const SageLayerArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int COUT = kLayerCout;
  constexpr int LPR = CIN / 4;                         // lanes per row: 4 or 5
  constexpr int GPW = 32 / LPR;                        // row groups per warp: 8 or 6
  constexpr int NWARPS = kLayerThreads / 32;
  constexpr int G = NWARPS * GPW;                      // rows in flight per pass of the CTA
  constexpr int NW = W / 2;
  constexpr int NOWN = COUT / LPR;                     // output columns per lane: 5 or 4
  constexpr bool EXTRA = NOWN == 5;                    // 4 lanes: the fifth column is 16 + sl
  static_assert(CIN == 16 || CIN == 20, "instantiated for 16 or 20 input channels");
  constexpr uint32_t row_bytes = CIN * 4u;

  const int nmax = a.nmax, parts = a.parts;
  const int nrow8 = (nmax + 7) & ~7;
  const int g = blockIdx.x / parts, p = blockIdx.x - g * parts;
  const int r0 = p * a.rpp;
  const int nrows = min(a.rpp, nrow8 - r0);            // > 0 by construction of parts, multiple of 8
  pdl_launch_dependents();
  const int64_t grow0 = (int64_t)g * nmax;
  const int64_t trow0 = (int64_t)g * nrow8 + r0;
  float* s_x = reinterpret_cast<float*>(smem_raw);
  // the CTA's slice of the index table (and weights / scales) behind the feature tile: brought in by bulk copies
  // at kernel start, so no global load sits between two rows of the gather loop
  unsigned char* s_idx = smem_raw + a.tile_smem;
  float* s_scale = reinterpret_cast<float*>(s_idx + (size_t)a.rpp * W * 2);
  float* s_wgt = s_scale + a.rpp;
  __shared__ __align__(8) uint64_t bar, bar_t;
  __shared__ float s_red[NWARPS][2][COUT];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gw = lane / LPR, sl = lane - gw * LPR;
  const bool active = gw < GPW;                        // CIN = 20: lanes 30, 31 idle
  const int grp = warp * GPW + (active ? gw : 0);
  const int lbase = lane - sl;                         // first lane of this row group
  const bool has_scale = a.ell_scale != nullptr;

  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_init(&bar_t, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (STRIDED) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
  }
  if (!STRIDED && tid < 2 * CIN) s_x[(size_t)nmax * CIN + tid] = 0.f;

  auto issue_tables = [&]() {                          // thread 0
    const uint32_t ib = (uint32_t)nrows * W * 2u, sb = has_scale ? (uint32_t)nrows * 4u : 0u;
    const uint32_t wb = WEIGHTED ? (uint32_t)nrows * W * 4u : 0u;
    mbar_expect_tx(&bar_t, ib + sb + wb);
    bulk_g2s(s_idx, a.ell + trow0 * W, ib, &bar_t);
    if (has_scale) bulk_g2s(s_scale, a.ell_scale + trow0, sb, &bar_t);
    if (WEIGHTED) bulk_g2s(s_wgt, a.ellw + trow0 * W, wb, &bar_t);
  };
  // W columns of this lane, in registers for the whole kernel: wr[k][0..3] = W[k][4sl..4sl+3], wr[k][4] = W[k][16+sl]
  float wr[CIN][NOWN];
  float br[NOWN];
  auto load_weights = [&]() {
#pragma unroll
    for (int k = 0; k < CIN; ++k) {
      const float4 v = ldg_f4(a.w + k * COUT + 4 * sl);
      wr[k][0] = v.x; wr[k][1] = v.y; wr[k][2] = v.z; wr[k][3] = v.w;
      if (EXTRA) wr[k][NOWN - 1] = __ldg(a.w + k * COUT + 16 + sl);
    }
    if (a.b != nullptr) {
      const float4 v = ldg_f4(a.b + 4 * sl);
      br[0] = v.x; br[1] = v.y; br[2] = v.z; br[3] = v.w;
      if (EXTRA) br[NOWN - 1] = __ldg(a.b + 16 + sl);
    } else {
#pragma unroll
      for (int j = 0; j < NOWN; ++j) br[j] = 0.f;
    }
  };

  __syncthreads();                                     // barriers initialised before anything is issued on them
  if (a.early) {
    if (tid == 0) issue_tables();
    load_weights();
  }
  pdl_wait();
  if (tid == 0) {
    if (STRIDED) {
      const int nboxes = (nmax + a.box_rows - 1) / a.box_rows;
      mbar_expect_tx(&bar, (uint32_t)nboxes * (uint32_t)a.box_rows * row_bytes);
      for (int bx = 0; bx < nboxes; ++bx)
        tma_load_2d(smem_raw + (size_t)bx * a.box_rows * row_bytes, &map_x, 0, (int)grow0 + bx * a.box_rows, &bar);
    } else {
      const uint32_t tile_bytes = (uint32_t)nmax * row_bytes;
      mbar_expect_tx(&bar, tile_bytes);
      const unsigned char* src = reinterpret_cast<const unsigned char*>(a.x + grow0 * CIN);
      for (uint32_t off = 0; off < tile_bytes; off += 32768u) {
        const uint32_t n = min(32768u, tile_bytes - off);
        bulk_g2s(smem_raw + off, src + off, n, &bar);
      }
    }
    if (!a.early) issue_tables();
  }
  if (!a.early) load_weights();
  const int nn = a.num_nodes ? __ldg(a.num_nodes + g) : nmax;   // real nodes of this graph
  mbar_wait(&bar_t, 0);
  mbar_wait(&bar, 0);
  if (STRIDED) {
    if (tid < 2 * CIN) s_x[(size_t)nmax * CIN + tid] = 0.f;
    __syncthreads();
  }

  const uint32_t lane_base = smem_u32(smem_raw) + (uint32_t)sl * 16u;
  const uint32_t idx_base = smem_u32(s_idx);
  const uint32_t ovf_mark = (uint32_t)nmax + 1u;
  float s1[NOWN], s2[NOWN];
#pragma unroll
  for (int j = 0; j < NOWN; ++j) { s1[j] = 0.f; s2[j] = 0.f; }

  auto load_idx = [&](uint32_t* dst, int t) {          // the row's W slots: 16 or 32 bytes, broadcast to its lanes
    const bool on = active && t < nrows;
#pragma unroll
    for (int q = 0; q < NW / 4; ++q) {
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (on) asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                           : "r"(idx_base + (uint32_t)t * (W * 2u) + q * 16u));
      dst[4 * q] = v.x; dst[4 * q + 1] = v.y; dst[4 * q + 2] = v.z; dst[4 * q + 3] = v.w;
    }
  };
  uint32_t cur[NW], nxt[NW];
  load_idx(cur, grp);
  for (int t0 = 0; t0 < nrows; t0 += G) {              // warp-uniform trip count: the shuffles need every lane
    const int t = t0 + grp;
    load_idx(nxt, t + G);
    const int lrow = ((r0 + t) & ~7) + ell_served_row(cur[0]);
    const bool row_on = active && t < nrows && lrow < nmax;
    const int64_t row = grow0 + (row_on ? lrow : 0);
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row_on) {
      float s = has_scale ? s_scale[t] : 1.0f;
      if ((cur[0] & kEllColMask) != ovf_mark) {
        float wg[WEIGHTED ? W : 1];
        if (WEIGHTED) {
#pragma unroll
          for (int q = 0; q < W / 4; ++q) {
            const float4 v = *reinterpret_cast<const float4*>(s_wgt + (size_t)t * W + 4 * q);
            wg[WEIGHTED ? 4 * q : 0] = v.x; wg[WEIGHTED ? 4 * q + 1 : 0] = v.y;
            wg[WEIGHTED ? 4 * q + 2 : 0] = v.z; wg[WEIGHTED ? 4 * q + 3 : 0] = v.w;
          }
        }
        F32x4 acc;
        acc.lo = 0ull; acc.hi = 0ull;
        int cnt = 0;
#pragma unroll
        for (int u = 0; u < W; ++u) {
          const uint32_t j = ell_slot(cur, u);
          if (!WEIGHTED) cnt += (j != (uint32_t)nmax) ? 1 : 0;
          const F32x4 v = lds_x4(lane_base + j * row_bytes);
          if (WEIGHTED) fma_x4(acc, wg[WEIGHTED ? u : 0], v);
          else add_x4(acc, v);
        }
        if (!WEIGHTED && !has_scale) s = 1.0f / fmaxf((float)cnt, 1.0f);      // the DenseSAGEConv mean
        o = unpack_x4(acc);
      } else {
        const int eb = __ldg(a.rowptr + row), ee = __ldg(a.rowptr + row + 1);
        if (!WEIGHTED && !has_scale) s = 1.0f / fmaxf((float)(ee - eb), 1.0f);
        for (int k = eb; k < ee; ++k) {
          const uint32_t j = (uint32_t)(__ldg(a.col + k) - (int)grow0);
          const float wv = WEIGHTED ? __ldg(a.val + k) : 1.0f;
          const float4 v = lds_f4(lane_base + j * row_bytes);
          o.x = fmaf(wv, v.x, o.x); o.y = fmaf(wv, v.y, o.y); o.z = fmaf(wv, v.z, o.z); o.w = fmaf(wv, v.w, o.w);
        }
      }
      o.x *= s; o.y *= s; o.z *= s; o.w *= s;
      *reinterpret_cast<float4*>(a.sa + row * CIN + sl * 4) = o;
    }
    // all-gather of the aggregated row over its LPR lanes
    float full[CIN];
#pragma unroll
    for (int d = 0; d < LPR; ++d) {
      full[4 * d + 0] = __shfl_sync(0xffffffffu, o.x, lbase + d);
      full[4 * d + 1] = __shfl_sync(0xffffffffu, o.y, lbase + d);
      full[4 * d + 2] = __shfl_sync(0xffffffffu, o.z, lbase + d);
      full[4 * d + 3] = __shfl_sync(0xffffffffu, o.w, lbase + d);
    }
    float z[NOWN];
#pragma unroll
    for (int j = 0; j < NOWN; ++j) z[j] = br[j];
#pragma unroll
    for (int k = 0; k < CIN; ++k)
#pragma unroll
      for (int j = 0; j < NOWN; ++j) z[j] = fmaf(full[k], wr[k][j], z[j]);
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < NOWN; ++j) ss = fmaf(z[j], z[j], ss);
    float tot = 0.f;
#pragma unroll
    for (int d = 0; d < LPR; ++d) tot += __shfl_sync(0xffffffffu, ss, lbase + d);
    const float inv = 1.0f / fmaxf(sqrtf(tot), kL2NormEps);
    const bool valid = lrow < nn;
    float yv[NOWN];
#pragma unroll
    for (int j = 0; j < NOWN; ++j) yv[j] = valid ? act_fwd(a.act, z[j] * inv) : 0.f;
    if (row_on) {
      float* yrow = a.y + row * a.ldy;
      *reinterpret_cast<float4*>(yrow + 4 * sl) = make_float4(yv[0], yv[1], yv[2], yv[3]);
      if (EXTRA) yrow[16 + sl] = yv[NOWN - 1];
#pragma unroll
      for (int j = 0; j < NOWN; ++j) { s1[j] += yv[j]; s2[j] = fmaf(yv[j], yv[j], s2[j]); }
    }
#pragma unroll
    for (int q = 0; q < NW; ++q) cur[q] = nxt[q];
  }

  if (a.partial == nullptr) return;
  // column sums: over the warp's row groups (lanes with equal sl), then over the warps in double precision
#pragma unroll
  for (int j = 0; j < NOWN; ++j) {
    float t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int d = 0; d < GPW; ++d) {
      t1 += __shfl_sync(0xffffffffu, s1[j], sl + d * LPR);
      t2 += __shfl_sync(0xffffffffu, s2[j], sl + d * LPR);
    }
    if (gw == 0) {
      const int c = (EXTRA && j == NOWN - 1) ? 16 + sl : 4 * sl + j;
      s_red[warp][0][c] = t1;
      s_red[warp][1][c] = t2;
    }
  }
  __syncthreads();
  if (tid < 2 * COUT) {
    const int which = tid / COUT, c = tid - which * COUT;
    double v = 0.0;
#pragma unroll
    for (int wv = 0; wv < NWARPS; ++wv) v += (double)s_red[wv][which][c];
    a.partial[((int64_t)blockIdx.x * 2 + which) * COUT + c] = v;
  }
  // The CTA that finishes last turns the partials into mean / invstd / running statistics (fixed summation order:
  // the result does not depend on which CTA that is), so the layer needs no separate finalize launch.
  __shared__ int s_last;
  __shared__ double s_tot[2][COUT];
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(a.counter, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (tid < 2 * COUT * 8) {                            // 8 threads per (sum kind, column): 320 threads = 10 warps
    const int col = tid >> 3, sub = tid & 7;
    const int which = col / COUT, c = col - which * COUT;
    double v = 0.0;
    const double* src = a.partial + (int64_t)which * COUT + c;
    for (int bk = sub; bk < (int)gridDim.x; bk += 64) {        // 8 independent loads in flight, summed in order
      double t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) t[u] = (bk + 8 * u < (int)gridDim.x) ? __ldcg(src + (int64_t)(bk + 8 * u) * 2 * COUT) : 0.0;
#pragma unroll
      for (int u = 0; u < 8; ++u) v += t[u];
    }
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    if (sub == 0) s_tot[which][c] = v;
  }
  __syncthreads();
  if (tid < COUT) {                                    // same arithmetic as k_bn_finalize (colops.cu)
    const double rows = (double)gridDim.x / parts * nmax;
    const double mu = s_tot[0][tid] / rows;
    double var = s_tot[1][tid] / rows - mu * mu;
    if (var < 0.0) var = 0.0;
    a.mean[tid] = (float)mu;
    a.invstd[tid] = (float)(1.0 / sqrt(var + (double)a.eps));
    if (a.running_mean) a.running_mean[tid] = (1.0f - a.momentum) * a.running_mean[tid] + a.momentum * (float)mu;
    if (a.running_var) {
      const double unbiased = rows > 1.0 ? var * rows / (rows - 1.0) : var;
      a.running_var[tid] = (1.0f - a.momentum) * a.running_var[tid] + a.momentum * (float)unbiased;
    }
  }
  if (tid == 0) *a.counter = 0u;
}

template <int CIN, int W, bool WEIGHTED, bool STRIDED>
static int launch_layer(SageLayerArgs a, int num_graphs, int64_t ldx, bool pdl, void* work, int64_t work_bytes,
                        cudaStream_t st) {
  const int nmax = a.nmax;
  size_t smem = ((size_t)nmax + 2) * CIN * 4;
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  if (STRIDED) {
    int64_t rows_buf = 0;
    a.box_rows = pick_box_rows(nmax, CIN * 4, 1, 227 * 1024 - 4096, &rows_buf);
    CGC_REQUIRE(a.box_rows > 0, CGC_ERR_UNSUPPORTED, "cgc_sage_layer_ell_fwd: %d rows do not fit in shared memory", nmax);
    smem = (size_t)rows_buf * CIN * 4;
    CGC_REQUIRE(tensor_map_2d(&map, a.x, (uint64_t)CIN, (uint64_t)num_graphs * nmax, ldx, (uint32_t)CIN,
                              (uint32_t)a.box_rows),
                CGC_ERR_UNSUPPORTED, "cgc_sage_layer_ell_fwd: cuTensorMapEncodeTiled failed");
  }
  int parts = kNumSMs / num_graphs;
  if (parts < 1) parts = 1;
  const int nrow8 = (nmax + 7) & ~7;
  int rpp = (int)ceil_div(nrow8, parts);
  rpp = (rpp + 7) & ~7;
  // the table slice of the CTA's rows (indices, scales, weights) sits behind the tile: split finer if it does not fit
  a.tile_smem = (int)((smem + 127) & ~(size_t)127);
  const size_t per_row = (size_t)W * 2 + 4 + (WEIGHTED ? (size_t)W * 4 : 0);
  const size_t room = 227 * 1024 - 3072 - 128 - (size_t)a.tile_smem;      // 3 KB: static shared memory + slack
  if ((size_t)rpp * per_row > room) rpp = (int)(room / per_row) & ~7;
  CGC_REQUIRE(rpp >= 8, CGC_ERR_UNSUPPORTED, "cgc_sage_layer_ell_fwd: no shared memory left for the index table (nmax %d)", nmax);
  parts = (int)ceil_div(nrow8, rpp);
  a.parts = parts;
  a.rpp = rpp;
  smem = (size_t)a.tile_smem + (size_t)rpp * per_row;
  const int grid = num_graphs * parts;
  if (a.partial != nullptr) {
    CGC_REQUIRE((int64_t)grid * 2 * kLayerCout * 8 + kCounterBytes <= work_bytes, CGC_ERR_WORKSPACE,
                "cgc_sage_layer_ell_fwd: workspace too small");
    a.counter = reinterpret_cast<unsigned int*>(static_cast<unsigned char*>(work) + work_bytes - kCounterBytes);
  }
  auto kern = k_sage_layer_ell<CIN, W, WEIGHTED, STRIDED>;
  static size_t smem_set[kMaxDevices] = {};             // per instantiation and device
  CGC_CUDA(ensure_dynamic_smem(kern, smem, smem_set));
  PdlLaunch L(dim3((unsigned)grid), dim3(kLayerThreads), smem, st, pdl);
  CGC_CUDA(cudaLaunchKernelEx(&L.cfg, kern, map, a));
  CGC_CHECK_LAUNCH("k_sage_layer_ell");
  return 0;
}

}  // namespace cgc

using namespace cgc;

extern "C" int cgc_sage_layer_ell_supported(int32_t cin, int32_t cout, int32_t nmax) {
  if (cout != kLayerCout || (cin != 16 && cin != 20) || nmax <= 0 || nmax + 1 > (int)kEllColMask) return 0;
  int64_t rows_buf = 0;
  // the strided variant needs the larger buffer (box overhang); require it so both input layouts are served
  return pick_box_rows(nmax, cin * 4, 1, 227 * 1024 - 4096 - 8 * 104, &rows_buf) > 0 ? 1 : 0;   // tile + >= 8 table rows
}

extern "C" int cgc_sage_layer_ell_fwd(const uint16_t* ell, const float* ell_val, const float* ell_scale,
                                      const uint32_t* perm, int32_t width, const int32_t* rowptr, const int32_t* col,
                                      const float* val, const float* x, int64_t ldx, int32_t num_graphs, int32_t nmax,
                                      int32_t cin, const float* w, const float* b, const int32_t* num_nodes,
                                      int32_t cout, int act, float* sa, float* y, int64_t ldy, float* running_mean,
                                      float* running_var, float momentum, float eps, float* mean, float* invstd,
                                      void* work, int64_t work_bytes, int flags, void* stream) {
  CGC_REQUIRE(ell && perm && rowptr && col && x && w && sa && y && num_graphs > 0 && nmax > 0, CGC_ERR_ARG,
              "cgc_sage_layer_ell_fwd: bad args");
  CGC_REQUIRE(ell_scale != nullptr || ell_val == nullptr, CGC_ERR_ARG,
              "cgc_sage_layer_ell_fwd: ell_scale may be NULL (= the mean 1 / max(deg, 1)) only for unweighted tables");
  CGC_REQUIRE(cgc_sage_layer_ell_supported(cin, cout, nmax), CGC_ERR_UNSUPPORTED,
              "cgc_sage_layer_ell_fwd: instantiated for cin 16 / 20 and cout 20 with the tile in shared memory "
              "(got cin %d cout %d nmax %d)", cin, cout, nmax);
  CGC_REQUIRE(width == 8 || width == 16, CGC_ERR_UNSUPPORTED, "cgc_sage_layer_ell_fwd: width must be 8 or 16");
  CGC_REQUIRE(ldx >= cin && ldx % 4 == 0 && ldy >= cout && ldy % 4 == 0 && aligned16(x) && aligned16(y) &&
                  aligned16(sa) && aligned16(ell) && aligned16(w) && (b == nullptr || aligned16(b)),
              CGC_ERR_UNSUPPORTED, "cgc_sage_layer_ell_fwd: ldx / ldy multiples of 4 and 16B-aligned pointers needed");
  CGC_REQUIRE((ell_val == nullptr) == (val == nullptr), CGC_ERR_ARG,
              "cgc_sage_layer_ell_fwd: val and ell_val must be given together");
  CGC_REQUIRE(ell_val == nullptr || aligned16(ell_val), CGC_ERR_UNSUPPORTED,
              "cgc_sage_layer_ell_fwd: ell_val must be 16B-aligned");
  CGC_REQUIRE((flags & CGC_GATHER_ACCUMULATE) == 0, CGC_ERR_ARG, "cgc_sage_layer_ell_fwd: ACCUMULATE is not a layer flag");
  CGC_REQUIRE(mean == nullptr || (invstd != nullptr && work != nullptr), CGC_ERR_ARG,
              "cgc_sage_layer_ell_fwd: statistics need invstd and a workspace");
  SageLayerArgs a;
  memset(&a, 0, sizeof(a));
  a.ell = ell; a.ellw = ell_val; a.ell_scale = ell_scale; a.perm = perm; a.rowptr = rowptr; a.col = col; a.val = val;
  a.x = x; a.w = w; a.b = b; a.num_nodes = num_nodes; a.sa = sa; a.y = y; a.ldy = ldy;
  a.partial = mean ? reinterpret_cast<double*>(work) : nullptr;
  a.running_mean = running_mean; a.running_var = running_var; a.mean = mean; a.invstd = invstd;
  a.momentum = momentum; a.eps = eps;
  a.nmax = nmax; a.act = act;
  a.early = (flags & CGC_GATHER_TABLES_READY) ? 1 : 0;
  cudaStream_t st = as_stream(stream);
  const bool wt = ell_val != nullptr, strided = ldx != cin;
  const bool pdl = (flags & CGC_GATHER_NO_PDL) == 0;
  if (!pdl) a.early = 0;
#define CGC_LAYER(CIN_, W_, WT_)                                                                                   \
  do {                                                                                                             \
    if (strided)                                                                                                   \
      return launch_layer<CIN_, W_, WT_, true>(a, num_graphs, ldx, pdl, work, work_bytes, st);                        \
    return launch_layer<CIN_, W_, WT_, false>(a, num_graphs, ldx, pdl, work, work_bytes, st);                         \
  } while (0)
  if (cin == 16) {
    if (width == 8) { if (wt) CGC_LAYER(16, 8, true); else CGC_LAYER(16, 8, false); }
    else            { if (wt) CGC_LAYER(16, 16, true); else CGC_LAYER(16, 16, false); }
  } else {
    if (width == 8) { if (wt) CGC_LAYER(20, 8, true); else CGC_LAYER(20, 8, false); }
    else            { if (wt) CGC_LAYER(20, 16, true); else CGC_LAYER(20, 16, false); }
  }
#undef CGC_LAYER
}
