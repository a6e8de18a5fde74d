// Sparse neighbour gather over the padded-row CSR/CSC.
//
//   cgc_spmm                 generic: any feature width / leading dimension, gathers through L2.
//   cgc_sage_gather_mean_fwd the SAGE roofline kernel: one graph's feature tile is staged in shared
//                            memory with bulk async copies (TMA, cp.async.bulk + mbarrier), column
//                            indices are staged with coalesced loads, neighbour rows are then read
//                            from shared memory and rows are written back coalesced.
//
// Reference arithmetic: torch.matmul(adj, x) / adj.sum(-1).clamp(min=1) of PyG 1.2.1
// DenseSAGEConv.forward (call sites model/network.py:114-116) and adj @ s of _diff_pool (:207).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "ell_gather.cuh"

namespace cgc {

// ------------------------------------------------------------------ generic gather
template <int VEC>
struct Vec;
template <>
struct Vec<4> {
  using T = float4;
  static __device__ __forceinline__ T zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  static __device__ __forceinline__ T load(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
  static __device__ __forceinline__ void fma(T& a, float w, const T& v) {
    a.x = fmaf(w, v.x, a.x); a.y = fmaf(w, v.y, a.y); a.z = fmaf(w, v.z, a.z); a.w = fmaf(w, v.w, a.w);
  }
  static __device__ __forceinline__ void store(float* p, T v, float s, bool acc) {
    float4* q = reinterpret_cast<float4*>(p);
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    if (acc) { float4 o = *q; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
    *q = v;
  }
};
template <>
struct Vec<1> {
  using T = float;
  static __device__ __forceinline__ T zero() { return 0.f; }
  static __device__ __forceinline__ T load(const float* p) { return __ldg(p); }
  static __device__ __forceinline__ void fma(T& a, float w, const T& v) { a = fmaf(w, v, a); }
  static __device__ __forceinline__ void store(float* p, T v, float s, bool acc) {
    v *= s;
    if (acc) v += *p;
    *p = v;
  }
};

// L lanes cooperate on one row (L = power of two <= 32); each lane owns chunks sl, sl+L, ... of the
// row, kept CH at a time in registers while the row's edges are streamed.
template <int VEC, int CH>
__global__ void __launch_bounds__(256) k_spmm(const int* __restrict__ rowptr, const int* __restrict__ col,
                                              const float* __restrict__ val, int64_t rows,
                                              const float* __restrict__ x, int64_t ldx, int nchunks,
                                              const float* __restrict__ src_scale,
                                              const float* __restrict__ row_scale, float* __restrict__ out,
                                              int64_t ldo, int accumulate, int L) {
  using V = Vec<VEC>;
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t r = t / L;
  int sl = (int)(t - r * L);
  if (r >= rows) return;
  const int b = __ldg(rowptr + r), e = __ldg(rowptr + r + 1);
  const float rsc = row_scale ? __ldg(row_scale + r) : 1.0f;
  for (int c0 = sl; c0 < nchunks; c0 += L * CH) {
    typename V::T acc[CH];
#pragma unroll
    for (int u = 0; u < CH; ++u) acc[u] = V::zero();
#pragma unroll 4
    for (int i = b; i < e; ++i) {
      const int j = __ldg(col + i);
      float w = val ? __ldg(val + i) : 1.0f;
      if (src_scale) w *= __ldg(src_scale + j);
      const float* xr = x + (int64_t)j * ldx;
#pragma unroll
      for (int u = 0; u < CH; ++u) {
        int c = c0 + u * L;
        if (c < nchunks) V::fma(acc[u], w, V::load(xr + (int64_t)c * VEC));
      }
    }
    float* orow = out + r * ldo;
#pragma unroll
    for (int u = 0; u < CH; ++u) {
      int c = c0 + u * L;
      if (c < nchunks) V::store(orow + (int64_t)c * VEC, acc[u], rsc, accumulate != 0);
    }
  }
}

// ------------------------------------------------------------------ TMA-staged SAGE gather
constexpr int kGatherThreads = 512;
constexpr int kGatherMaxRows = 1024;        // rows of one CTA (rowptr staging buffer)
constexpr int kStageUnroll = 16;            // independent loads in flight per thread while staging
constexpr uint32_t kBulkChunk = 32768;      // bytes per cp.async.bulk

// coalesced, deeply unrolled global -> shared staging of n 4-byte words
template <typename T>
__device__ __forceinline__ void stage_words(T* __restrict__ dst, const T* __restrict__ src, int n) {
  for (int base = 0; base < n; base += kStageUnroll * kGatherThreads) {
    T tmp[kStageUnroll];
#pragma unroll
    for (int u = 0; u < kStageUnroll; ++u) {
      int i = base + u * kGatherThreads + threadIdx.x;
      if (i < n) tmp[u] = __ldg(src + i);
    }
#pragma unroll
    for (int u = 0; u < kStageUnroll; ++u) {
      int i = base + u * kGatherThreads + threadIdx.x;
      if (i < n) dst[i] = tmp[u];
    }
  }
}

// grid = num_graphs * parts.  CTA (g, p) owns rows [p*rpp, min((p+1)*rpp, nmax)) of graph g.
// LPR lanes per row, each lane one float4 of the feature row.
template <int LPR>
__global__ void __launch_bounds__(kGatherThreads, 1)
k_sage_gather_tile(const int* __restrict__ rowptr, const int* __restrict__ col, const float* __restrict__ val,
                   const float* __restrict__ row_scale, const float* __restrict__ x, int nmax, int feat,
                   int parts, int rpp, int col_cap, float* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int g = blockIdx.x / parts, p = blockIdx.x - g * parts;
  const int r0 = p * rpp;
  const int nrows = min(rpp, nmax - r0);
  if (nrows <= 0) return;
  const int64_t grow0 = (int64_t)g * nmax;           // first padded row of this graph
  const uint32_t tile_bytes = (uint32_t)nmax * feat * 4u;

  float* s_x = reinterpret_cast<float*>(smem_raw);
  int* s_ptr = reinterpret_cast<int*>(smem_raw + tile_bytes);
  int* s_col = s_ptr + (kGatherMaxRows + 4);
  float* s_val = reinterpret_cast<float*>(s_col + col_cap);
  __shared__ __align__(8) uint64_t bar;

  const int* gptr = rowptr + grow0 + r0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(&bar, tile_bytes);
    const unsigned char* src = reinterpret_cast<const unsigned char*>(x + grow0 * feat);
    for (uint32_t off = 0; off < tile_bytes; off += kBulkChunk) {
      uint32_t n = min(kBulkChunk, tile_bytes - off);
      bulk_g2s(smem_raw + off, src + off, n, &bar);
    }
  }
  // While the tile is in flight: the two row pointers bounding this CTA's edges (broadcast loads, in
  // flight together with the rowptr slice), then the column indices.  No barrier until everything landed.
  const int e0 = __ldg(gptr);
  const int nnz = __ldg(gptr + nrows) - e0;
  stage_words(s_ptr, gptr, nrows + 1);
  const bool staged = nnz <= col_cap;
  if (staged) {
    stage_words(s_col, col + e0, nnz);
    if (val != nullptr) stage_words(s_val, val + e0, nnz);
  }
  __syncthreads();
  mbar_wait(&bar, 0);

  const int sl = threadIdx.x % LPR;                  // float4 chunk of the row
  const int grp = threadIdx.x / LPR;
  constexpr int kGroups = kGatherThreads / LPR;
  const bool lane_on = sl * 4 < feat;
  const int base_col = (int)grow0;
  // The groups of one warp read the column lists of consecutive rows; starting the upper half of the
  // groups 4 edges later keeps their smem words in different banks (deg 8: 8 groups x 8 words wrap at 32).
  const int rot_req = ((grp % (32 / LPR)) >= (16 / LPR)) ? 4 : 0;
  for (int r = grp; r < nrows; r += kGroups) {
    const int b = s_ptr[r] - e0, e = s_ptr[r + 1] - e0;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane_on) {
      const int deg = e - b;
      const int rot = deg > rot_req ? rot_req : 0;
#pragma unroll 4
      for (int t = 0; t < deg; ++t) {
        int tt = t + rot;
        if (tt >= deg) tt -= deg;
        const int i = b + tt;
        int j;
        float w = 1.0f;
        if (staged) {
          j = s_col[i] - base_col;
          if (val != nullptr) w = s_val[i];
        } else {
          j = __ldg(col + e0 + i) - base_col;
          if (val != nullptr) w = __ldg(val + e0 + i);
        }
        const float4 v = *reinterpret_cast<const float4*>(s_x + (size_t)j * feat + sl * 4);
        acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y);
        acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
      }
      float s;
      if (row_scale != nullptr) s = __ldg(row_scale + grow0 + r0 + r);
      else if (val == nullptr) s = 1.0f / fmaxf((float)deg, 1.0f);
      else s = 1.0f;
      acc.x *= s; acc.y *= s; acc.z *= s; acc.w *= s;
      *reinterpret_cast<float4*>(out + (grow0 + r0 + r) * feat + sl * 4) = acc;
    }
  }
}


// ------------------------------------------------------------------ ELL16 gather (fixed row pitch)
// Row-padded CSR ("ELL16"): every padded row owns W slots of uint16 LOCAL column ids (column - g*nmax);
// unused slots hold `nmax`, which addresses an all-zero row appended to the shared-memory tile.  A row whose
// degree exceeds W is marked by slot 0 == nmax + 1 (a second zero row) and walked through the CSR instead.  The
// fixed pitch removes the rowptr -> column-index dependency: a row's whole index list is ONE 16-byte load whose
// address is known at kernel start, so index loads fly together with the TMA tile copy and nothing but the
// feature tile is staged through shared memory.
//
// Layout: graph g owns nrow8 = round_up(nmax, 8) table rows; table row g*nrow8 + 8*blk + p is the row served by
// lane group p of the warp that handles block blk, and perm[g*nrow8/8 + blk] (3 bits per position) names the
// local row 8*blk + ((perm >> 3p) & 7) it belongs to.
//
// Neighbour rows are read from the shared-memory tile, where a 64-byte feature row covers half of the 32 banks,
// chosen by the parity of its index: the two rows served by one quarter-warp (positions 2k, 2k+1) collide iff
// their current neighbours have equal parity.  With `pair_rows` the builder therefore
//   * orders the slots "even columns first" on even positions and "odd columns first" on odd positions,
//   * sorts the 8 rows of a block by their number e of even columns and pairs smallest with largest, so that
//     e_A + e_B is as close to W as the block allows: the pair then collides in only |e_A + e_B - W| slots.
// (Fetching one neighbour per row straight from global memory during the tile copy was tried and lost: the
// dependent 64-byte reads go to DRAM at random and delay the tile barrier by 2 us.)
// Only the summation order inside a row changes; the adjacency is exactly that of to_dense_adj (utils.py:29-33).
constexpr int kEllMaxW = 16;
// Slot 0 of every table row also carries, in bits 12-14, the local row q (within the 8-row block) the table row
// serves -- the same information as the `perm` words, readable without a second load.  Column ids therefore use 12
// bits: nmax + 1 <= 0xFFF (every kernel that gathers from a shared-memory tile needs nmax <= ~3600 anyway).

__global__ void __launch_bounds__(128) k_csr_to_ell16(const int* __restrict__ rowptr, const int* __restrict__ col,
                                                      const float* __restrict__ val,
                                                      const float* __restrict__ row_scale, int num_graphs, int nmax,
                                                      int W, int pair_rows, uint16_t* __restrict__ ell,
                                                      float* __restrict__ ellw, float* __restrict__ ell_scale,
                                                      uint32_t* __restrict__ perm) {
  const int nrow8 = (nmax + 7) & ~7;
  const int nb8 = nrow8 >> 3;
  const int64_t blk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (blk >= (int64_t)num_graphs * nb8) return;
  const int g = (int)(blk / nb8), lb = (int)(blk - (int64_t)g * nb8);
  const int base = g * nmax;
  const int n = W;
  const int pad_par = nmax & 1;
  int deg[8], key[8];
  for (int q = 0; q < 8; ++q) {
    const int lr = lb * 8 + q;
    deg[q] = -1;                      // row does not exist (beyond nmax)
    key[q] = n / 2;
    if (lr >= nmax) continue;
    const int64_t r = (int64_t)base + lr;
    const int b = __ldg(rowptr + r), e = __ldg(rowptr + r + 1);
    deg[q] = e - b;
    if (deg[q] > W) continue;         // overflow row
    if (deg[q] == 0) { key[q] = pad_par == 0 ? n : 0; continue; }
    int ev = 0;
    for (int i = b; i < e; ++i) ev += ((__ldg(col + i) - base) & 1) == 0;
    const int pads = W - deg[q];
    key[q] = ev + (pad_par == 0 ? pads : 0);
  }
  int order[8] = {0, 1, 2, 3, 4, 5, 6, 7};
  if (pair_rows) {
    for (int i = 1; i < 8; ++i) {     // stable insertion sort by key
      const int o = order[i];
      int j = i - 1;
      while (j >= 0 && key[order[j]] > key[o]) { order[j + 1] = order[j]; --j; }
      order[j + 1] = o;
    }
  }
  uint32_t pw = 0;
  for (int p = 0; p < 8; ++p) {
    const int q = pair_rows ? ((p & 1) ? order[7 - (p >> 1)] : order[p >> 1]) : p;
    pw |= (uint32_t)q << (3 * p);
    uint16_t* o = ell + (blk * 8 + p) * W;
    float* ow = ellw ? ellw + (blk * 8 + p) * W : nullptr;
    // the row's output scale, in table order: row_scale[row], else the mean of the binary adjacency (DenseSAGEConv)
    float sc = 1.0f;
    if (deg[q] >= 0) {
      if (row_scale) sc = __ldg(row_scale + base + lb * 8 + q);
      else if (!val) sc = 1.0f / fmaxf((float)deg[q], 1.0f);
    }
    ell_scale[blk * 8 + p] = sc;
    int pos = 0;
    if (deg[q] > W) {
      o[pos] = (uint16_t)(nmax + 1);
      if (ow) ow[pos] = 0.f;
      ++pos;
    } else if (deg[q] > 0) {
      const int64_t r = (int64_t)base + lb * 8 + q;
      const int b = __ldg(rowptr + r), e = b + deg[q];
      const int npass = pair_rows ? 2 : 1;
      for (int pass = 0; pass < npass; ++pass) {
        const int want = pass == 0 ? (p & 1) : 1 - (p & 1);
        for (int i = b; i < e; ++i) {
          const int c = __ldg(col + i) - base;
          if (pair_rows && (c & 1) != want) continue;
          if (ow) ow[pos] = __ldg(val + i);
          o[pos++] = (uint16_t)c;
        }
      }
    }
    for (; pos < W; ++pos) {
      o[pos] = (uint16_t)nmax;
      if (ow) ow[pos] = 0.f;
    }
    o[0] = (uint16_t)(o[0] | (q << kEllRowShift));     // the local row (within its 8-row block) this table row serves
  }
  perm[blk] = pw;
}

// grid = num_graphs * parts; CTA (g, p) owns table rows [p*rpp, min((p+1)*rpp, nrow8)) of graph g (rpp % 8 == 0).
// LPR lanes per row, one float4 of the feature row per lane; group grp = tid / LPR serves table rows grp, grp+G, ...
// Before the tile barrier every thread loads the index vectors (weights, row scales) of its RB rows (next batch:
// prefetched); after it the slot loop is pure LDS.128 + packed FADD2 / FFMA2 on registers.
// Everything a thread reads before the barrier (indices, weights, scale, position map) is addressed by the table
// row alone: no load depends on another.
//
// STRIDED: x has a leading dimension > feat (a column slice of the concat buffer, layers 2-3 of a GNN_Module): the
// tile is staged by 2-D TMA boxes {feat, box_rows} described by map_x instead of 1-D bulk copies; out always
// takes ldo.  ACC: out += result (the CSC gathers of the backward accumulate into slices of the gradient buffer).
//
// Programmatic dependent launch: the kernel is launched with programmatic stream serialization, so its prologue
// (barrier init, zero rows and -- when the caller says the tables were not written by the kernel right before it,
// `early_tables` -- all index / scale / position-map loads) runs under the tail of the previous kernel;
// pdl_wait() then orders the tile copy (x is the previous kernel's output) and every store after it.
struct EllGatherArgs {
  const uint16_t* ell;
  const float* ellw;
  const float* ell_scale;
  const uint32_t* perm;
  const int* rowptr;
  const int* col;
  const float* val;
  const float* x;
  float* out;
  int64_t ldo;
  int nmax, feat, parts, rpp;
  int n_hi, rpp_hi;      // balanced split: the first n_hi graphs get parts_hi CTAs of rpp_hi rows each, so that the
  int parts_hi;          // grid covers every SM (64 graphs: 20 x 3 + 44 x 2 = 148 CTAs; clusters of 2: 10 x 4 + 54 x 2)
  int box_rows;          // STRIDED: rows per TMA box
  int early_tables;      // index tables may be read before pdl_wait()
  int lab_stage;         // CGC_LAB builds: 1 = exit at entry, 2 = exit after the tile copy (timeline measurements)
  uint32_t chunk;        // bytes per 1-D bulk copy
};

// CTA -> (graph, first table row, rows): graphs [0, n_hi) are cut into parts + 1 pieces of rpp_hi rows, the rest into
// `parts` pieces of rpp rows
__device__ __forceinline__ void ell_cta_rows(const EllGatherArgs& a, int nrow8, int& g, int& r0, int& nrows) {
  const int hi_ctas = a.n_hi * a.parts_hi;
  int p, rpp;
  if ((int)blockIdx.x < hi_ctas) {
    g = blockIdx.x / a.parts_hi;
    p = blockIdx.x - g * a.parts_hi;
    rpp = a.rpp_hi;
  } else {
    const int b2 = blockIdx.x - hi_ctas;
    g = a.n_hi + b2 / a.parts;
    p = b2 - (b2 / a.parts) * a.parts;
    rpp = a.rpp;
  }
  r0 = p * rpp;
  nrows = min(rpp, nrow8 - r0);
}

// Unweighted tables may come without a scale vector (ell_scale == NULL): the row's scale is then the DenseSAGEConv
// mean 1 / max(deg, 1) with deg counted from the row's own slots; the served row comes from the slot-0 tag.
// (The BASELINE configuration under a programmatic launch takes k_sage_gather_ell_lean below instead.)
template <int W, int LPR, bool WEIGHTED, int RB, int THREADS, bool STRIDED, bool ACC>
__global__ void __launch_bounds__(THREADS, 1)
k_sage_gather_ell(const __grid_constant__ CUtensorMap map_x, const EllGatherArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int G = THREADS / LPR;
  constexpr int NW = W / 2;                            // 32-bit index words per row
  const int nmax = a.nmax, feat = a.feat, parts = a.parts;
  const int nrow8 = (nmax + 7) & ~7;
  int g, r0, nrows;
  ell_cta_rows(a, nrow8, g, r0, nrows);
  const int p = r0 / max(nrows, 1);                    // (lab stage 3 only)
  (void)p;
  pdl_launch_dependents();                             // the next PDL kernel may start its prologue
#ifdef CGC_LAB
  if (a.lab_stage == 1) return;
#endif
  if (nrows <= 0) return;
  const int64_t grow0 = (int64_t)g * nmax;             // first padded row of the graph (x / out / rowptr)
  const int64_t trow0 = (int64_t)g * nrow8 + r0;       // first table row of this CTA
  const uint32_t row_bytes = (uint32_t)feat * 4u;
  const uint32_t tile_bytes = (uint32_t)nmax * row_bytes;
  float* s_x = reinterpret_cast<float*>(smem_raw);
  __shared__ __align__(8) uint64_t bar;

  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (STRIDED) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
  }
  // two zero rows behind the tile: index nmax = empty slot, nmax + 1 = overflow mark (STRIDED: the overhang of the
  // last TMA box lands there, so they are written after the barrier instead)
  if (!STRIDED && threadIdx.x < 2 * feat) s_x[(size_t)nmax * feat + threadIdx.x] = 0.f;

  const int sl = threadIdx.x % LPR;
  const int grp = threadIdx.x / LPR;
  const bool lane_on = sl * 4 < feat;
  const uint32_t lane_base = smem_u32(smem_raw) + (lane_on ? sl : 0) * 16u;
  const uint32_t pad_word = (uint32_t)nmax | ((uint32_t)nmax << 16);
  const uint32_t ovf_mark = (uint32_t)nmax + 1u;

  uint32_t idx[RB][NW];
  float wgt[WEIGHTED ? RB : 1][WEIGHTED ? W : 1];
  float rsc[RB];
  auto load_batch = [&](int i0) {
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      const int t = (i0 + i) * G + grp;                // table row inside the CTA
      const bool on = t < nrows;
      const int64_t trow = trow0 + (on ? t : 0);
      rsc[i] = (on && a.ell_scale != nullptr) ? __ldg(a.ell_scale + trow) : 1.0f;
#pragma unroll
      for (int q = 0; q < NW / 4; ++q) {
        uint4 v = make_uint4(pad_word, pad_word, pad_word, pad_word);
        if (on) v = ldg_u4(a.ell + trow * W + q * 8);
        idx[i][4 * q] = v.x; idx[i][4 * q + 1] = v.y; idx[i][4 * q + 2] = v.z; idx[i][4 * q + 3] = v.w;
      }
      if (WEIGHTED) {
#pragma unroll
        for (int q = 0; q < W / 4; ++q) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (on) v = ldg_f4(a.ellw + trow * W + q * 4);
          wgt[WEIGHTED ? i : 0][WEIGHTED ? 4 * q : 0] = v.x;
          wgt[WEIGHTED ? i : 0][WEIGHTED ? 4 * q + 1 : 0] = v.y;
          wgt[WEIGHTED ? i : 0][WEIGHTED ? 4 * q + 2 : 0] = v.z;
          wgt[WEIGHTED ? i : 0][WEIGHTED ? 4 * q + 3 : 0] = v.w;
        }
      }
    }
  };

  if (a.early_tables) load_batch(0);                   // flies under the previous kernel's tail
  pdl_wait();                                          // x (and, without early_tables, the tables) are now final
  if (threadIdx.x == 0) {
    if (STRIDED) {
      const int nboxes = (nmax + a.box_rows - 1) / a.box_rows;
      mbar_expect_tx(&bar, (uint32_t)nboxes * (uint32_t)a.box_rows * row_bytes);
      for (int bx = 0; bx < nboxes; ++bx)
        tma_load_2d(smem_raw + (size_t)bx * a.box_rows * row_bytes, &map_x, 0, (int)grow0 + bx * a.box_rows, &bar);
    } else {
      uint32_t lo = 0, hi = tile_bytes;
#ifdef CGC_LAB
      if (a.lab_stage == 3) {                          // lab: each CTA copies only its share of the tile
        const uint32_t share = (tile_bytes / parts) & ~15u;
        lo = p * share; hi = lo + share;
      } else if (a.lab_stage >= 4) {                   // lab: every CTA copies the first 1/(stage-2) of the tile
        hi = (tile_bytes / (a.lab_stage - 2)) & ~15u;
      }
#endif
      mbar_expect_tx(&bar, hi - lo);
      const unsigned char* src = reinterpret_cast<const unsigned char*>(a.x + grow0 * feat);
      for (uint32_t off = lo; off < hi; off += a.chunk) {
        uint32_t n = min(a.chunk, hi - off);
        bulk_g2s(smem_raw + off, src + off, n, &bar);
      }
    }
  }
  if (!a.early_tables) load_batch(0);
  __syncthreads();                                     // zero rows visible, barrier initialised
  mbar_wait(&bar, 0);
  if (STRIDED) {
    if (threadIdx.x < 2 * feat) s_x[(size_t)nmax * feat + threadIdx.x] = 0.f;
    __syncthreads();
  }
#ifdef CGC_LAB
  if (a.lab_stage >= 2) return;
#endif

  for (int i0 = 0; i0 * G < nrows; i0 += RB) {
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      const int t = (i0 + i) * G + grp;
      if (t >= nrows) continue;                        // whole quarter-warps at a time
      const int lrow = ((r0 + t) & ~7) + ell_served_row(idx[i][0]);
      if (lrow >= nmax) continue;                      // padding of the graph's last 8-row block
      const int64_t row = grow0 + lrow;
      float4 o;
      float s = rsc[i];
      const bool mean_scale = !WEIGHTED && a.ell_scale == nullptr;            // scale = 1 / max(deg, 1)
      if ((idx[i][0] & kEllColMask) != ovf_mark) {
        F32x4 acc;
        acc.lo = 0ull; acc.hi = 0ull;
        int cnt = 0;
#pragma unroll
        for (int u = 0; u < W; ++u) {
          const uint32_t j = ell_slot(idx[i], u);
          if (!WEIGHTED) cnt += (j != (uint32_t)nmax) ? 1 : 0;
          const F32x4 v = lds_x4(lane_base + j * row_bytes);
          if (WEIGHTED) fma_x4(acc, wgt[WEIGHTED ? i : 0][WEIGHTED ? u : 0], v);
          else add_x4(acc, v);
        }
        if (mean_scale) s = 1.0f / fmaxf((float)cnt, 1.0f);
        o = unpack_x4(acc);
      } else {
        // degree > W: walk the CSR (column ids are global padded rows)
        const int b = __ldg(a.rowptr + row), e = __ldg(a.rowptr + row + 1);
        if (mean_scale) s = 1.0f / fmaxf((float)(e - b), 1.0f);
        o = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = b; k < e; ++k) {
          const uint32_t j = (uint32_t)(__ldg(a.col + k) - (int)grow0);
          const float w = WEIGHTED ? __ldg(a.val + k) : 1.0f;
          const float4 v = lds_f4(lane_base + j * row_bytes);
          o.x = fmaf(w, v.x, o.x); o.y = fmaf(w, v.y, o.y); o.z = fmaf(w, v.z, o.z); o.w = fmaf(w, v.w, o.w);
        }
      }
      if (lane_on) {
        o.x *= s; o.y *= s; o.z *= s; o.w *= s;
        float4* dst = reinterpret_cast<float4*>(a.out + row * a.ldo + sl * 4);
        if (ACC) {
          const float4 old = *dst;
          o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
        }
        *dst = o;
      }
    }
    if ((i0 + RB) * G < nrows) load_batch(i0 + RB);
  }
}


// ------------------------------------------------------------------ ELL16 gather, two CTAs per SM
// The roofline configuration (pitch 8, unweighted, 16 features, contiguous operands) in a shape that lets TWO CTAs
// share an SM: 512 threads x 8 rows, 64 registers (a row costs its 4 index registers and nothing else: served row from
// the slot-0 tag, scale = 1 / max(deg, 1) counted from the slots), 2 x 113.6 KB of shared memory.  Under
// programmatic dependent launch the NEXT launch's CTAs are then resident, with their index vectors loaded, while
// this launch still gathers, and pull their tile into L2 before griddepcontrol.wait.  It only pays on a grid that
// covers every SM (ell_row_split): with 128 CTAs the block scheduler doubles the next launch's CTAs up on the 20 idle
// SMs and the launch waits for those (measured 8.7 us against 6.0 - 6.2 us).
constexpr int kLeanThreads = 512;
constexpr uint32_t kLeanChunk = 16384;   // bytes per bulk copy of the tile (8 KB and up measure the same, 4 KB and
                                         // below are slower: profiles/gather_chunk_r2.txt)
constexpr int kLeanRows = 8;

// CL > 1: the CL CTAs of a graph form a thread-block cluster and share ONE read of the tile: each issues every CL-th
// chunk as a multicast bulk copy that lands in all of them (L2 -> SM traffic of a launch 17 MB -> 7.4 MB at B = 64).
template <int CL>
__global__ void __launch_bounds__(kLeanThreads, 2)
k_sage_gather_ell_lean(const EllGatherArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int W = 8, NW = 4, LPR = 4, G = kLeanThreads / LPR, RB = kLeanRows;
  const int nmax = a.nmax, feat = a.feat;
  const int nrow8 = (nmax + 7) & ~7;
  int g, r0, nrows;
  ell_cta_rows(a, nrow8, g, r0, nrows);
  pdl_launch_dependents();
  const int64_t grow0 = (int64_t)g * nmax;
  const int64_t trow0 = (int64_t)g * nrow8 + r0;
  const uint32_t row_bytes = (uint32_t)feat * 4u;
  const uint32_t tile_bytes = (uint32_t)nmax * row_bytes;
  float* s_x = reinterpret_cast<float*>(smem_raw);
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 2 * feat) s_x[(size_t)nmax * feat + threadIdx.x] = 0.f;
  uint32_t crank = 0;
  if (CL > 1) {
    crank = cluster_rank();
    // arrive now, wait only right before the first multicast copy is issued: every CTA's barrier exists before a
    // peer's copy counts on it, and nobody idles here for a peer that was scheduled later
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  }
  __syncthreads();                                       // barrier and pad rows visible to the CTA (still in the prologue)
  const int sl = threadIdx.x % LPR, grp = threadIdx.x / LPR;
  const uint32_t lane_base = smem_u32(smem_raw) + sl * 16u;
  const uint32_t pad_word = (uint32_t)nmax | ((uint32_t)nmax << 16);
  const uint32_t ovf_mark = (uint32_t)nmax + 1u;
  uint32_t idx[RB][NW];
  auto load_all = [&]() {
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      const int t = i * G + grp;
      uint4 v = make_uint4(pad_word, pad_word, pad_word, pad_word);
      if (t < nrows) v = ldg_u4(a.ell + (trow0 + t) * W);
      idx[i][0] = v.x; idx[i][1] = v.y; idx[i][2] = v.z; idx[i][3] = v.w;
    }
  };
  if (a.early_tables) load_all();
  // Pull the graph's feature tile into L2 while the previous kernel is still running.  This is safe whatever that
  // kernel does to x: L2 is the coherence point of the device (a later write updates the same L2 line) and the bulk
  // copy below reads through L2 after griddepcontrol.wait; it only moves the HBM round trip of the tile under the
  // previous launch's gather phase, when the memory system is idle.  In the training step x is L2-resident already.
  if (threadIdx.x == 32) {
    const unsigned char* src = reinterpret_cast<const unsigned char*>(a.x + grow0 * feat);
    for (uint32_t off = crank * a.chunk; off < tile_bytes; off += CL * a.chunk) {
      const uint32_t n = min(a.chunk, tile_bytes - off);
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + off), "r"(n) : "memory");
    }
  }
  pdl_wait();
  if (CL > 1) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, tile_bytes);                    // the whole tile: this CTA's chunks and its peers'
    const unsigned char* src = reinterpret_cast<const unsigned char*>(a.x + grow0 * feat);
    for (uint32_t off = crank * a.chunk; off < tile_bytes; off += CL * a.chunk) {
      const uint32_t n = min(a.chunk, tile_bytes - off);
      if (CL > 1) bulk_g2s_multicast(smem_raw + off, src + off, n, &bar, (uint16_t)((1u << CL) - 1u));
      else bulk_g2s(smem_raw + off, src + off, n, &bar);
    }
  }
  if (!a.early_tables) load_all();
  __syncthreads();
  mbar_wait(&bar, 0);
#pragma unroll
  for (int i = 0; i < RB; ++i) {
    const int t = i * G + grp;
    if (t >= nrows) continue;
    const int lrow = ((r0 + t) & ~7) + ell_served_row(idx[i][0]);
    if (lrow >= nmax) continue;
    const int64_t row = grow0 + lrow;
    float4 o;
    float s;
    if ((idx[i][0] & kEllColMask) != ovf_mark) {
      F32x4 acc;
      acc.lo = 0ull; acc.hi = 0ull;
      int cnt = 0;
#pragma unroll
      for (int u = 0; u < W; ++u) {
        const uint32_t j = ell_slot(idx[i], u);
        cnt += (j != (uint32_t)nmax) ? 1 : 0;
        add_x4(acc, lds_x4(lane_base + j * row_bytes));
      }
      s = 1.0f / fmaxf((float)cnt, 1.0f);
      o = unpack_x4(acc);
    } else {
      const int b = __ldg(a.rowptr + row), e = __ldg(a.rowptr + row + 1);
      s = 1.0f / fmaxf((float)(e - b), 1.0f);
      o = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int k = b; k < e; ++k) {
        const float4 v = lds_f4(lane_base + (uint32_t)(__ldg(a.col + k) - (int)grow0) * row_bytes);
        o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w;
      }
    }
    o.x *= s; o.y *= s; o.z *= s; o.w *= s;
    *reinterpret_cast<float4*>(a.out + row * a.ldo + sl * 4) = o;
  }
  // (no cluster barrier at exit: every CTA has waited for its whole tile, i.e. for every copy that targets ITS shared
  //  memory; a copy this CTA issued completes in the peers without it)
}


// ------------------------------------------------------------------ ELL16 gather, table slice staged in shared memory
// Same result as k_sage_gather_ell for the table shapes whose rows do not fit one register batch per thread (pitch 16
// and / or weighted): there a thread serves its rows in up to 7 batches, each of which started with global loads of
// indices and weights whose latency nothing covered (12 us instead of 6 on cold tables).  Here the CTA's slice of
// the table -- indices, scales, weights -- is brought into shared memory behind the feature tile by three bulk copies
// at kernel start (under the previous kernel's tail when the tables are ready), and the row loop reads it with
// broadcast LDS, the next row's indices one iteration ahead.
struct EllStagedArgs {
  EllGatherArgs g;
  int tile_smem;          // bytes of the feature tile buffer (the table slice follows it)
};

template <int W, int LPR, bool WEIGHTED, int THREADS, bool STRIDED, bool ACC>
__global__ void __launch_bounds__(THREADS, 1)
k_sage_gather_ell_staged(const __grid_constant__ CUtensorMap map_x, const EllStagedArgs sa) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const EllGatherArgs& a = sa.g;
  constexpr int G = THREADS / LPR;
  constexpr int NW = W / 2;
  const int nmax = a.nmax, feat = a.feat;
  const int nrow8 = (nmax + 7) & ~7;
  int g, r0, nrows;
  ell_cta_rows(a, nrow8, g, r0, nrows);                // > 0 rows by construction, multiple of 8
  pdl_launch_dependents();
  const int64_t grow0 = (int64_t)g * nmax;
  const int64_t trow0 = (int64_t)g * nrow8 + r0;
  const uint32_t row_bytes = (uint32_t)feat * 4u;
  float* s_x = reinterpret_cast<float*>(smem_raw);
  unsigned char* s_idx = smem_raw + sa.tile_smem;
  float* s_scale = reinterpret_cast<float*>(s_idx + (size_t)a.rpp * W * 2);
  float* s_wgt = s_scale + a.rpp;
  __shared__ __align__(8) uint64_t bar, bar_t;
  const bool has_scale = a.ell_scale != nullptr;
  const int tid = threadIdx.x;

  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_init(&bar_t, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (STRIDED) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
  }
  if (!STRIDED && tid < 2 * feat) s_x[(size_t)nmax * feat + tid] = 0.f;
  auto issue_tables = [&]() {                          // thread 0
    const uint32_t ib = (uint32_t)nrows * W * 2u, sb = has_scale ? (uint32_t)nrows * 4u : 0u;
    const uint32_t wb = WEIGHTED ? (uint32_t)nrows * W * 4u : 0u;
    mbar_expect_tx(&bar_t, ib + sb + wb);
    bulk_g2s(s_idx, a.ell + trow0 * W, ib, &bar_t);
    if (has_scale) bulk_g2s(s_scale, a.ell_scale + trow0, sb, &bar_t);
    if (WEIGHTED) bulk_g2s(s_wgt, a.ellw + trow0 * W, wb, &bar_t);
  };
  __syncthreads();
  if (a.early_tables && tid == 0) issue_tables();
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
      const unsigned char* src = reinterpret_cast<const unsigned char*>(a.x + grow0 * feat);
      for (uint32_t off = 0; off < tile_bytes; off += a.chunk) {
        const uint32_t n = min(a.chunk, tile_bytes - off);
        bulk_g2s(smem_raw + off, src + off, n, &bar);
      }
    }
    if (!a.early_tables) issue_tables();
  }
  mbar_wait(&bar_t, 0);
  mbar_wait(&bar, 0);
  if (STRIDED) {
    if (tid < 2 * feat) s_x[(size_t)nmax * feat + tid] = 0.f;
    __syncthreads();
  }

  const int sl = tid % LPR, grp = tid / LPR;
  const bool lane_on = sl * 4 < feat;
  const uint32_t lane_base = smem_u32(smem_raw) + (lane_on ? sl : 0) * 16u;
  const uint32_t idx_base = smem_u32(s_idx);
  const uint32_t ovf_mark = (uint32_t)nmax + 1u;
  auto load_idx = [&](uint32_t* dst, int t) {
#pragma unroll
    for (int q = 0; q < NW / 4; ++q) {
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (t < nrows) asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                                  : "r"(idx_base + (uint32_t)t * (W * 2u) + q * 16u));
      dst[4 * q] = v.x; dst[4 * q + 1] = v.y; dst[4 * q + 2] = v.z; dst[4 * q + 3] = v.w;
    }
  };
  uint32_t cur[NW], nxt[NW];
  load_idx(cur, grp);
  for (int t = grp; t < nrows; t += G) {
    load_idx(nxt, t + G);
    const int lrow = ((r0 + t) & ~7) + ell_served_row(cur[0]);
    if (lrow < nmax) {
      const int64_t row = grow0 + lrow;
      float s = has_scale ? s_scale[t] : 1.0f;
      float4 o;
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
        if (!WEIGHTED && !has_scale) s = 1.0f / fmaxf((float)cnt, 1.0f);
        o = unpack_x4(acc);
      } else {
        const int b = __ldg(a.rowptr + row), e = __ldg(a.rowptr + row + 1);
        if (!WEIGHTED && !has_scale) s = 1.0f / fmaxf((float)(e - b), 1.0f);
        o = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = b; k < e; ++k) {
          const uint32_t j = (uint32_t)(__ldg(a.col + k) - (int)grow0);
          const float w = WEIGHTED ? __ldg(a.val + k) : 1.0f;
          const float4 v = lds_f4(lane_base + j * row_bytes);
          o.x = fmaf(w, v.x, o.x); o.y = fmaf(w, v.y, o.y); o.z = fmaf(w, v.z, o.z); o.w = fmaf(w, v.w, o.w);
        }
      }
      if (lane_on) {
        o.x *= s; o.y *= s; o.z *= s; o.w *= s;
        float4* dst = reinterpret_cast<float4*>(a.out + row * a.ldo + sl * 4);
        if (ACC) {
          const float4 old = *dst;
          o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
        }
        *dst = o;
      }
    }
#pragma unroll
    for (int q = 0; q < NW; ++q) cur[q] = nxt[q];
  }
}


// ------------------------------------------------------------------ wide gather over the ELL16 table
// out[r, :] (+)= scale[r] * sum_e val[e] * x[col[e], :] for wide feature matrices (A*S of _diff_pool, F = 1140, and
// its transpose in the backward).  The generic k_spmm reads every x row once per referencing edge through L2
// (8x 525 MB at the BASELINE shape: L2-bandwidth bound, 425 us); here a work item is (graph, 16-column slice): the
// slice of the graph's rows -- a strided [nmax, 16] tile -- is brought into shared memory ONCE by 2-D TMA
// (cp.async.bulk.tensor, boxes of 16 columns x `box_rows` rows), the rows are gathered from shared memory exactly
// like in k_sage_gather_ell, and the CTA (persistent, one per SM) double-buffers: the tile of item i+1 lands
// while item i is gathered.
constexpr int kWideThreads = 512;
constexpr int kWideCols = 16;                // columns per slice (64-byte shared-memory rows, as in the F = 16 gather)

template <int W, bool WEIGHTED, int RB>                  // RB = rows per thread and batch
__global__ void __launch_bounds__(kWideThreads, 1)
k_gather_wide_ell(const __grid_constant__ CUtensorMap map_x, const uint16_t* __restrict__ ell,
                  const float* __restrict__ ellw, const float* __restrict__ ell_scale,
                  const uint32_t* __restrict__ perm, const int* __restrict__ rowptr, const int* __restrict__ col,
                  const float* __restrict__ val, int num_graphs, int nmax, int feat, int box_rows, int nbuf,
                  float* __restrict__ out, int64_t ldo, int accumulate) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int G = kWideThreads / 4;
  constexpr int NW = W / 2;
  const int nrow8 = (nmax + 7) & ~7;
  const int nboxes = (nmax + box_rows - 1) / box_rows;
  const uint32_t buf_rows = (uint32_t)max(nboxes * box_rows, nmax + 2);          // tile (+ TMA overhang), >= 2 zero rows
  const uint32_t buf_bytes = buf_rows * 64u;
  const uint32_t tile_tx = (uint32_t)nboxes * (uint32_t)box_rows * 64u;
  const int nslices = (feat + kWideCols - 1) / kWideCols;
  const int64_t nitems = (int64_t)num_graphs * nslices;
  const int nbatch = (nrow8 + G * RB - 1) / (G * RB);      // batches per item
  __shared__ __align__(8) uint64_t full[2];
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
  }
  __syncthreads();
  auto issue = [&](int64_t item, int b) {                 // thread 0: all boxes of one tile onto full[b]
    const int g = (int)(item / nslices), sidx = (int)(item - (int64_t)g * nslices);
    mbar_expect_tx(&full[b], tile_tx);
    for (int bx = 0; bx < nboxes; ++bx)
      tma_load_2d(smem_raw + (size_t)b * buf_bytes + (size_t)bx * box_rows * 64, &map_x, sidx * kWideCols,
                  g * nmax + bx * box_rows, &full[b]);
  };
  const int64_t first = blockIdx.x, step = gridDim.x;
  if (tid == 0) {
    if (first < nitems) issue(first, 0);
    if (nbuf > 1 && first + step < nitems) issue(first + step, 1);
  }
  const int sl = tid & 3, grp = tid >> 2;
  const uint32_t pad_word = (uint32_t)nmax | ((uint32_t)nmax << 16);
  const uint32_t ovf_mark = (uint32_t)nmax + 1u;

  // Index vectors (weights, scale, position map) of one batch of RB rows.  They are loaded one batch AHEAD of
  // their use -- across item boundaries too -- so that their L2 latency hides behind the previous batch's gather.
  struct Batch {
    uint32_t idx[RB][NW];
    float wgt[WEIGHTED ? RB : 1][WEIGHTED ? W : 1];
    float rsc[RB];
    uint32_t pw[RB];
  };
  auto load_batch = [&](Batch& bt, int64_t item, int bi) {
    const int g = (int)(item / nslices);
    const int64_t trow0 = (int64_t)g * nrow8;
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      const int t = (bi * RB + i) * G + grp;
      const bool on = item < nitems && t < nrow8;
      const int64_t trow = trow0 + (on ? t : 0);
      bt.pw[i] = on ? __ldg(perm + (trow >> 3)) : 0xFFFFFFFFu;
      bt.rsc[i] = (on && ell_scale != nullptr) ? __ldg(ell_scale + trow) : 1.0f;
#pragma unroll
      for (int q = 0; q < NW / 4; ++q) {
        uint4 v = make_uint4(pad_word, pad_word, pad_word, pad_word);
        if (on) v = ldg_u4(ell + trow * W + q * 8);
        bt.idx[i][4 * q] = v.x; bt.idx[i][4 * q + 1] = v.y; bt.idx[i][4 * q + 2] = v.z; bt.idx[i][4 * q + 3] = v.w;
      }
      if (WEIGHTED) {
#pragma unroll
        for (int q = 0; q < W / 4; ++q) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (on) v = ldg_f4(ellw + trow * W + q * 4);
          bt.wgt[WEIGHTED ? i : 0][WEIGHTED ? 4 * q : 0] = v.x;
          bt.wgt[WEIGHTED ? i : 0][WEIGHTED ? 4 * q + 1 : 0] = v.y;
          bt.wgt[WEIGHTED ? i : 0][WEIGHTED ? 4 * q + 2 : 0] = v.z;
          bt.wgt[WEIGHTED ? i : 0][WEIGHTED ? 4 * q + 3 : 0] = v.w;
        }
      }
    }
  };

  uint32_t ph0 = 0u, ph1 = 0u;                            // barrier parities (scalars: no local-memory array)
  int it = 0;
  Batch cur, nxt;
  load_batch(cur, first, 0);
  for (int64_t item = first; item < nitems; item += step, ++it) {
    const int b = nbuf > 1 ? (it & 1) : 0;
    const int g = (int)(item / nslices), sidx = (int)(item - (int64_t)g * nslices);
    const int c0 = sidx * kWideCols + sl * 4;             // first output column of this lane
    const bool lane_on = c0 < feat;
    const int64_t grow0 = (int64_t)g * nmax;
    const uint32_t lane_base = smem_u32(smem_raw) + (uint32_t)b * buf_bytes + (uint32_t)sl * 16u;
    if (b == 0) { mbar_wait(&full[0], ph0); ph0 ^= 1u; }
    else { mbar_wait(&full[1], ph1); ph1 ^= 1u; }
    // zero rows for empty slots / overflow marks (the TMA overhang may have landed there)
    if (tid < 32) reinterpret_cast<float*>(smem_raw + (size_t)b * buf_bytes)[(size_t)nmax * 16 + tid] = 0.f;
    __syncthreads();
    for (int bi = 0; bi < nbatch; ++bi) {
      // prefetch the next batch (of this item, or the first one of the CTA's next item)
      if (bi + 1 < nbatch) load_batch(nxt, item, bi + 1);
      else load_batch(nxt, item + step, 0);
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        const int t = (bi * RB + i) * G + grp;
        if (t >= nrow8) continue;
        const int lrow = (t & ~7) + (int)((cur.pw[i] >> (3 * (t & 7))) & 7u);
        if (lrow >= nmax) continue;
        const int64_t row = grow0 + lrow;
        float4 o;
        if ((cur.idx[i][0] & kEllColMask) != ovf_mark) {
          F32x4 acc;
          acc.lo = 0ull; acc.hi = 0ull;
#pragma unroll
          for (int u = 0; u < W; ++u) {
            const uint32_t j = ell_slot(cur.idx[i], u);
            const F32x4 v = lds_x4(lane_base + j * 64u);
            if (WEIGHTED) fma_x4(acc, cur.wgt[WEIGHTED ? i : 0][WEIGHTED ? u : 0], v);
            else add_x4(acc, v);
          }
          o = unpack_x4(acc);
        } else {
          const int e0 = __ldg(rowptr + row), e1 = __ldg(rowptr + row + 1);
          o = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int k = e0; k < e1; ++k) {
            const uint32_t j = (uint32_t)(__ldg(col + k) - (int)grow0);
            const float w = WEIGHTED ? __ldg(val + k) : 1.0f;
            const float4 v = lds_f4(lane_base + j * 64u);
            o.x = fmaf(w, v.x, o.x); o.y = fmaf(w, v.y, o.y); o.z = fmaf(w, v.z, o.z); o.w = fmaf(w, v.w, o.w);
          }
        }
        if (lane_on) {
          const float s = cur.rsc[i];
          o.x *= s; o.y *= s; o.z *= s; o.w *= s;
          float4* dst = reinterpret_cast<float4*>(out + row * ldo + c0);
          if (accumulate) {
            const float4 old = *dst;
            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
          }
          *dst = o;
        }
      }
      cur = nxt;
    }
    __syncthreads();                                      // everyone is done with buffer b
    if (tid == 0) {
      const int64_t nx = item + (nbuf > 1 ? 2 : 1) * step;
      if (nx < nitems) issue(nx, b);
    }
  }
}

}  // namespace cgc

using namespace cgc;

static int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

extern "C" int cgc_spmm(const int32_t* rowptr, const int32_t* col, const float* val, int64_t rows,
                        const float* x, int64_t ldx, int32_t feat, const float* src_scale,
                        const float* row_scale, float* out, int64_t ldo, int accumulate, void* stream) {
  CGC_REQUIRE(rowptr && col && x && out && rows > 0 && feat > 0 && ldx >= feat && ldo >= feat, CGC_ERR_ARG,
              "cgc_spmm: bad args");
  cudaStream_t st = as_stream(stream);
  const bool vec = (feat % 4 == 0) && (ldx % 4 == 0) && (ldo % 4 == 0) && aligned16(x) && aligned16(out);
  if (vec) {
    int nchunks = feat / 4;
    int L = pow2_ceil(nchunks) > 32 ? 32 : pow2_ceil(nchunks);
    int64_t threads = rows * L;
    unsigned grid = (unsigned)ceil_div(threads, 256);
    if (nchunks > L)
      k_spmm<4, 4><<<grid, 256, 0, st>>>(rowptr, col, val, rows, x, ldx, nchunks, src_scale, row_scale, out, ldo,
                                         accumulate, L);
    else
      k_spmm<4, 1><<<grid, 256, 0, st>>>(rowptr, col, val, rows, x, ldx, nchunks, src_scale, row_scale, out, ldo,
                                         accumulate, L);
  } else {
    int L = pow2_ceil(feat) > 32 ? 32 : pow2_ceil(feat);
    int64_t threads = rows * L;
    unsigned grid = (unsigned)ceil_div(threads, 256);
    if (feat > L)
      k_spmm<1, 4><<<grid, 256, 0, st>>>(rowptr, col, val, rows, x, ldx, feat, src_scale, row_scale, out, ldo,
                                         accumulate, L);
    else
      k_spmm<1, 1><<<grid, 256, 0, st>>>(rowptr, col, val, rows, x, ldx, feat, src_scale, row_scale, out, ldo,
                                         accumulate, L);
  }
  CGC_CHECK_LAUNCH("k_spmm");
  return 0;
}

extern "C" int cgc_sage_gather_mean_fwd(const int32_t* rowptr, const int32_t* col, const float* val,
                                        const float* row_scale, const float* x, int32_t num_graphs,
                                        int32_t nmax, int32_t feat, float* out, void* stream) {
  CGC_REQUIRE(rowptr && col && x && out && num_graphs > 0 && nmax > 0 && feat > 0, CGC_ERR_ARG,
              "cgc_sage_gather_mean_fwd: bad args");
  CGC_REQUIRE(feat % 4 == 0 && feat <= 32 && aligned16(x) && aligned16(out), CGC_ERR_UNSUPPORTED,
              "cgc_sage_gather_mean_fwd: feat must be a multiple of 4 and <= 32 (got %d), 16B-aligned x/out",
              feat);
  const int64_t tile_bytes = (int64_t)nmax * feat * 4;
  const int64_t fixed = tile_bytes + (kGatherMaxRows + 4) * 4;
  const int64_t budget = 220 * 1024;
  CGC_REQUIRE(fixed + 4096 <= budget, CGC_ERR_UNSUPPORTED,
              "cgc_sage_gather_mean_fwd: feature tile of %lld B does not fit in shared memory",
              (long long)tile_bytes);
  int parts = kNumSMs / num_graphs;
  if (parts < 1) parts = 1;
  int min_parts = (int)ceil_div(nmax, kGatherMaxRows);
  if (parts < min_parts) parts = min_parts;
  int rpp = (int)ceil_div(nmax, parts);
  int64_t cap = (budget - fixed) / (val ? 8 : 4);
  if (cap > 16384) cap = 16384;
  cap &= ~(int64_t)511;
  size_t smem = (size_t)(fixed + cap * (val ? 8 : 4));
  cudaStream_t st = as_stream(stream);
  const int lpr = feat <= 16 ? 4 : 8;
  auto kern = lpr == 4 ? k_sage_gather_tile<4> : k_sage_gather_tile<8>;
  CGC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<num_graphs * parts, kGatherThreads, smem, st>>>(rowptr, col, val, row_scale, x, nmax, feat, parts, rpp,
                                                         (int)cap, out);
  CGC_CHECK_LAUNCH("k_sage_gather_tile");
  return 0;
}

extern "C" int64_t cgc_ell16_rows(int32_t num_graphs, int32_t nmax) {
  return (int64_t)num_graphs * ((nmax + 7) & ~7);
}

extern "C" int cgc_csr_to_ell16(const int32_t* rowptr, const int32_t* col, const float* val, const float* row_scale,
                                int32_t num_graphs, int32_t nmax, int32_t width, int pair_rows, uint16_t* ell,
                                float* ell_val, float* ell_scale, uint32_t* perm, void* stream) {
  CGC_REQUIRE(rowptr && col && ell && ell_scale && perm && num_graphs > 0 && nmax > 0, CGC_ERR_ARG,
              "cgc_csr_to_ell16: bad args");
  CGC_REQUIRE(width == 8 || width == 16, CGC_ERR_UNSUPPORTED, "cgc_csr_to_ell16: width must be 8 or 16 (got %d)",
              width);
  CGC_REQUIRE(nmax + 1 <= kEllColMask, CGC_ERR_UNSUPPORTED, "cgc_csr_to_ell16: nmax %d does not fit 12-bit local ids", nmax);
  CGC_REQUIRE((ell_val == nullptr) == (val == nullptr), CGC_ERR_ARG,
              "cgc_csr_to_ell16: val and ell_val must be given together");
  const int64_t blocks8 = cgc_ell16_rows(num_graphs, nmax) / 8;
  k_csr_to_ell16<<<(unsigned)ceil_div(blocks8, 128), 128, 0, as_stream(stream)>>>(rowptr, col, val, row_scale, num_graphs,
                                                                                  nmax, width, pair_rows, ell, ell_val,
                                                                                  ell_scale, perm);
  CGC_CHECK_LAUNCH("k_csr_to_ell16");
  return 0;
}

static int ell_row_split(EllGatherArgs& a, int num_graphs, int parts_override);

template <int W, int LPR, bool WEIGHTED, int RB, int THREADS, bool STRIDED, bool ACC>
static int launch_ell(EllGatherArgs a, int num_graphs, int64_t ldx, bool pdl, cudaStream_t st) {
  const int nmax = a.nmax, feat = a.feat;
  size_t smem = ((size_t)nmax + 2) * feat * 4;
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  if (STRIDED) {
    int64_t rows_buf = 0;
    a.box_rows = pick_box_rows(nmax, feat * 4, 1, 227 * 1024 - 1024, &rows_buf);
    CGC_REQUIRE(a.box_rows > 0, CGC_ERR_UNSUPPORTED,
                "cgc_sage_gather_ell: %d rows x %d features do not fit in shared memory", nmax, feat);
    smem = (size_t)rows_buf * feat * 4;
    CGC_REQUIRE(tensor_map_2d(&map, a.x, (uint64_t)feat, (uint64_t)num_graphs * nmax, ldx, (uint32_t)feat,
                              (uint32_t)a.box_rows),
                CGC_ERR_UNSUPPORTED, "cgc_sage_gather_ell: cuTensorMapEncodeTiled failed");
  }
  CGC_REQUIRE(smem + 1024 <= 227 * 1024, CGC_ERR_UNSUPPORTED,
              "cgc_sage_gather_ell: feature tile of %lld B does not fit in shared memory", (long long)smem);
  static const int parts_override = lab_env("CGC_GATHER_PARTS", 0);      // lab switches: CGC_LAB builds only
  static const int chunk_override = lab_env("CGC_GATHER_CHUNK", 0);
  static const int stage = lab_env("CGC_GATHER_STAGE", 0);
  const int grid = ell_row_split(a, num_graphs, parts_override);
  a.lab_stage = stage;
  a.chunk = chunk_override > 0 ? (uint32_t)chunk_override : kBulkChunk;
  auto kern = k_sage_gather_ell<W, LPR, WEIGHTED, RB, THREADS, STRIDED, ACC>;
  static size_t smem_set[kMaxDevices] = {};             // per instantiation and device
  CGC_CUDA(ensure_dynamic_smem(kern, smem, smem_set));
  PdlLaunch L(dim3((unsigned)grid), dim3(THREADS), smem, st, pdl);
  CGC_CUDA(cudaLaunchKernelEx(&L.cfg, kern, map, a));
  CGC_CHECK_LAUNCH("k_sage_gather_ell");
  return 0;
}

// Row split of the gather grids: `parts` CTAs per graph (rpp rows each) and, when the grid would leave SMs idle, one
// extra CTA for the first n_hi graphs (rpp_hi rows each).  Returns the grid size.
// cluster > 1: CTAs come in clusters of that many per graph (parts = cluster); the spare SMs then go to graphs that
// get TWO clusters (2 x cluster CTAs).
static int ell_row_split(EllGatherArgs& a, int num_graphs, int parts_override, int cluster) {
  const int nrow8 = (a.nmax + 7) & ~7;
  int parts = parts_override > 0 ? parts_override : kNumSMs / num_graphs;
  if (parts < 1) parts = 1;
  int rpp = (int)ceil_div(nrow8, parts);
  rpp = (rpp + 7) & ~7;                                 // CTAs own whole 8-row blocks of the table
  parts = (int)ceil_div(nrow8, rpp);
  a.parts = parts;
  a.rpp = rpp;
  a.n_hi = 0;
  a.rpp_hi = rpp;
  a.parts_hi = parts + 1;
  const int spare = kNumSMs - parts * num_graphs;       // SMs a plain split leaves idle
  if (cluster > 1) {
    a.parts_hi = 2 * parts;
    int rpp_hi = (int)ceil_div(nrow8, a.parts_hi);
    rpp_hi = (rpp_hi + 7) & ~7;
    if (parts == cluster && spare >= parts && (int)ceil_div(nrow8, rpp_hi) == a.parts_hi) {
      a.n_hi = spare / parts < num_graphs ? spare / parts : num_graphs;
      a.rpp_hi = rpp_hi;
    }
  } else if (parts_override <= 0 && spare > 0) {
    int rpp_hi = (int)ceil_div(nrow8, parts + 1);
    rpp_hi = (rpp_hi + 7) & ~7;
    if ((int)ceil_div(nrow8, rpp_hi) == parts + 1) {
      a.n_hi = spare < num_graphs ? spare : num_graphs;
      a.rpp_hi = rpp_hi;
    }
  }
  return a.n_hi * a.parts_hi + (num_graphs - a.n_hi) * parts;
}
static int ell_row_split(EllGatherArgs& a, int num_graphs, int parts_override) {
  return ell_row_split(a, num_graphs, parts_override, 1);
}

// The two-CTAs-per-SM variant of the roofline configuration (see k_sage_gather_ell_lean)
template <int CL>
static int launch_ell_lean_cl(const EllGatherArgs& a, int grid, size_t smem, bool pdl, cudaStream_t st) {
  auto kern = k_sage_gather_ell_lean<CL>;
  static size_t smem_set[kMaxDevices] = {};             // per instantiation and device
  CGC_CUDA(ensure_dynamic_smem(kern, smem, smem_set));
  static bool carve[kMaxDevices] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < kMaxDevices && !carve[dev]) {   // two CTAs per SM need the whole 228 KB as shared memory
    CGC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  (int)cudaSharedmemCarveoutMaxShared));
    carve[dev] = true;
  }
  PdlLaunch L(dim3((unsigned)grid), dim3(kLeanThreads), smem, st, pdl, CL);
  CGC_CUDA(cudaLaunchKernelEx(&L.cfg, kern, a));
  CGC_CHECK_LAUNCH("k_sage_gather_ell_lean");
  return 0;
}

static int launch_ell_lean(EllGatherArgs a, int num_graphs, bool pdl, bool multicast, cudaStream_t st) {
  const size_t smem = ((size_t)a.nmax + 2) * a.feat * 4;
  a.lab_stage = 0;
  static const int chunk_override = lab_env("CGC_GATHER_CHUNK", 0);
  a.chunk = chunk_override > 0 ? (uint32_t)chunk_override : kLeanChunk;
  // 2 or 4 CTAs per graph: one cluster per graph, the tile is read from L2 once and multicast (6.03 us against
  // 6.24 us for every CTA copying the whole tile itself, CGC_GATHER_NO_MULTICAST)
  EllGatherArgs c = a;
  const int want = kNumSMs / num_graphs;
  const int cl = want >= 4 ? 4 : want;
  if (multicast && (cl == 2 || cl == 4)) {
    const int grid = ell_row_split(c, num_graphs, cl, cl);
    if (c.parts == cl && grid % cl == 0)
      return cl == 2 ? launch_ell_lean_cl<2>(c, grid, smem, pdl, st) : launch_ell_lean_cl<4>(c, grid, smem, pdl, st);
  }
  const int grid = ell_row_split(a, num_graphs, 0);
  return launch_ell_lean_cl<1>(a, grid, smem, pdl, st);
}

// Launch of the staged-table variant; returns 1 ("not taken") when the table slice does not fit behind the tile.
template <int W, int LPR, bool WEIGHTED, int THREADS, bool STRIDED, bool ACC>
static int launch_ell_staged(EllGatherArgs a, int num_graphs, int64_t ldx, bool pdl, cudaStream_t st) {
  const int nmax = a.nmax, feat = a.feat;
  size_t tile = ((size_t)nmax + 2) * feat * 4;
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  if (STRIDED) {
    int64_t rows_buf = 0;
    a.box_rows = pick_box_rows(nmax, feat * 4, 1, 227 * 1024 - 2048, &rows_buf);
    if (a.box_rows <= 0) return 1;
    tile = (size_t)rows_buf * feat * 4;
    CGC_REQUIRE(tensor_map_2d(&map, a.x, (uint64_t)feat, (uint64_t)num_graphs * nmax, ldx, (uint32_t)feat,
                              (uint32_t)a.box_rows),
                CGC_ERR_UNSUPPORTED, "cgc_sage_gather_ell: cuTensorMapEncodeTiled failed");
  }
  EllStagedArgs sa;
  sa.tile_smem = (int)((tile + 127) & ~(size_t)127);
  const int grid = ell_row_split(a, num_graphs, 0);
  const size_t per_row = (size_t)W * 2 + 4 + (WEIGHTED ? (size_t)W * 4 : 0);
  const size_t smem = (size_t)sa.tile_smem + (size_t)a.rpp * per_row;
  if (smem + 1024 > 227 * 1024) return 1;
  a.lab_stage = 0;
  a.chunk = kBulkChunk;
  sa.g = a;
  auto kern = k_sage_gather_ell_staged<W, LPR, WEIGHTED, THREADS, STRIDED, ACC>;
  static size_t smem_set[kMaxDevices] = {};             // per instantiation and device
  CGC_CUDA(ensure_dynamic_smem(kern, smem, smem_set));
  PdlLaunch L(dim3((unsigned)grid), dim3(THREADS), smem, st, pdl);
  CGC_CUDA(cudaLaunchKernelEx(&L.cfg, kern, map, sa));
  CGC_CHECK_LAUNCH("k_sage_gather_ell_staged");
  return 0;
}

// flags of cgc_sage_gather_ell (include/cgc_sm100.h)
static int ell_gather_dispatch(const uint16_t* ell, const float* ell_val, const float* ell_scale, const uint32_t* perm,
                               int32_t width, const int32_t* rowptr, const int32_t* col, const float* val,
                               const float* x, int64_t ldx, int32_t num_graphs, int32_t nmax, int32_t feat, float* out,
                               int64_t ldo, int flags, cudaStream_t st) {
  EllGatherArgs a;
  memset(&a, 0, sizeof(a));
  a.ell = ell; a.ellw = ell_val; a.ell_scale = ell_scale; a.perm = perm;
  a.rowptr = rowptr; a.col = col; a.val = val; a.x = x; a.out = out; a.ldo = ldo;
  a.nmax = nmax; a.feat = feat;
  a.early_tables = (flags & CGC_GATHER_TABLES_READY) ? 1 : 0;
  const bool wt = ell_val != nullptr;
  const bool strided = ldx != feat;
  const bool acc = (flags & CGC_GATHER_ACCUMULATE) != 0;
  const bool pdl = (flags & CGC_GATHER_NO_PDL) == 0;
  if (!pdl) a.early_tables = 0;
  // Roofline configuration: the two-CTAs-per-SM variant under programmatic launch (the next launch's CTAs are
  // resident early and pull their tile into L2 while this launch gathers: 6.21 vs 6.51 us), the one-per-SM kernel
  // for a plain stream-ordered launch (7.74 vs 7.85 us: a tie), or as the flags say.
  const bool lean_ok = !wt && !strided && !acc && ell_scale == nullptr && width == 8 && feat == 16 && ldo == feat &&
                       2 * (((size_t)nmax + 2) * feat * 4 + 1024 + 128) <= 228 * 1024;
  if (lean_ok && !(flags & CGC_GATHER_ONE_PER_SM) && (pdl || (flags & CGC_GATHER_TWO_PER_SM)))
    return launch_ell_lean(a, num_graphs, pdl, (flags & CGC_GATHER_NO_MULTICAST) == 0, st);
  // pitch 16 and / or weighted tables: the variant with the table slice staged in shared memory, when it fits
#define CGC_ELL_STAGED(W_, LPR_, WT_)                                                                     \
  do {                                                                                                    \
    int rc_;                                                                                              \
    if (strided) rc_ = acc ? launch_ell_staged<W_, LPR_, WT_, 512, true, true>(a, num_graphs, ldx, pdl, st)     \
                           : launch_ell_staged<W_, LPR_, WT_, 512, true, false>(a, num_graphs, ldx, pdl, st);   \
    else rc_ = acc ? launch_ell_staged<W_, LPR_, WT_, 512, false, true>(a, num_graphs, ldx, pdl, st)            \
                   : launch_ell_staged<W_, LPR_, WT_, 512, false, false>(a, num_graphs, ldx, pdl, st);          \
    if (rc_ != 1) return rc_;                                                                             \
  } while (0)
  if (feat <= 16) {
    if (width == 8 && wt) CGC_ELL_STAGED(8, 4, true);
    if (width == 16 && wt) CGC_ELL_STAGED(16, 4, true);
    if (width == 16 && !wt) CGC_ELL_STAGED(16, 4, false);
  } else {
    if (width == 8 && wt) CGC_ELL_STAGED(8, 8, true);
    if (width == 16 && wt) CGC_ELL_STAGED(16, 8, true);
    if (width == 16 && !wt) CGC_ELL_STAGED(16, 8, false);
  }
#undef CGC_ELL_STAGED
#define CGC_ELL(W_, LPR_, WT_, RB_, TH_)                                                                  \
  do {                                                                                                    \
    if (strided) {                                                                                        \
      if (acc) return launch_ell<W_, LPR_, WT_, RB_, TH_, true, true>(a, num_graphs, ldx, pdl, st);            \
      return launch_ell<W_, LPR_, WT_, RB_, TH_, true, false>(a, num_graphs, ldx, pdl, st);                    \
    }                                                                                                     \
    if (acc) return launch_ell<W_, LPR_, WT_, RB_, TH_, false, true>(a, num_graphs, ldx, pdl, st);             \
    return launch_ell<W_, LPR_, WT_, RB_, TH_, false, false>(a, num_graphs, ldx, pdl, st);                     \
  } while (0)
  if (feat <= 16) {
    if (width == 8) {
      if (wt) CGC_ELL(8, 4, true, 2, 512);
      else CGC_ELL(8, 4, false, 4, 1024);
    } else {
      if (wt) CGC_ELL(16, 4, true, 1, 512); else CGC_ELL(16, 4, false, 4, 512);
    }
  } else {
    if (width == 8) { if (wt) CGC_ELL(8, 8, true, 2, 512); else CGC_ELL(8, 8, false, 4, 512); }
    else            { if (wt) CGC_ELL(16, 8, true, 1, 512); else CGC_ELL(16, 8, false, 4, 512); }
  }
#undef CGC_ELL
}

extern "C" int cgc_sage_gather_ell(const uint16_t* ell, const float* ell_val, const float* ell_scale,
                                   const uint32_t* perm, int32_t width, const int32_t* rowptr, const int32_t* col,
                                   const float* val, const float* x, int64_t ldx, int32_t num_graphs, int32_t nmax,
                                   int32_t feat, float* out, int64_t ldo, int flags, void* stream) {
  CGC_REQUIRE(ell && perm && rowptr && col && x && out && num_graphs > 0 && nmax > 0 && feat > 0, CGC_ERR_ARG,
              "cgc_sage_gather_ell: bad args");
  CGC_REQUIRE(ell_scale != nullptr || ell_val == nullptr, CGC_ERR_ARG,
              "cgc_sage_gather_ell: ell_scale may be NULL (= the mean 1 / max(deg, 1)) only for unweighted tables");
  CGC_REQUIRE(width == 8 || width == 16, CGC_ERR_UNSUPPORTED, "cgc_sage_gather_ell: width must be 8 or 16");
  CGC_REQUIRE(feat % 4 == 0 && feat <= 32 && ldx >= feat && ldo >= feat && ldx % 4 == 0 && ldo % 4 == 0 &&
                  aligned16(x) && aligned16(out) && aligned16(ell),
              CGC_ERR_UNSUPPORTED,
              "cgc_sage_gather_ell: feat must be a multiple of 4 and <= 32 (got %d), ldx / ldo multiples of 4, "
              "16B-aligned x/out/ell", feat);
  CGC_REQUIRE((ell_val == nullptr) == (val == nullptr), CGC_ERR_ARG,
              "cgc_sage_gather_ell: val and ell_val must be given together");
  CGC_REQUIRE(ell_val == nullptr || aligned16(ell_val), CGC_ERR_UNSUPPORTED,
              "cgc_sage_gather_ell: ell_val must be 16B-aligned");
  return ell_gather_dispatch(ell, ell_val, ell_scale, perm, width, rowptr, col, val, x, ldx, num_graphs, nmax, feat,
                             out, ldo, flags, as_stream(stream));
}

// contiguous x / out, plain store, tables possibly fresh: the round-1 entry point, kept for callers of that ABI
extern "C" int cgc_sage_gather_ell_fwd(const uint16_t* ell, const float* ell_val, const float* ell_scale,
                                       const uint32_t* perm, int32_t width, const int32_t* rowptr, const int32_t* col,
                                       const float* val, const float* x, int32_t num_graphs, int32_t nmax,
                                       int32_t feat, float* out, void* stream) {
  return cgc_sage_gather_ell(ell, ell_val, ell_scale, perm, width, rowptr, col, val, x, feat, num_graphs, nmax, feat,
                             out, feat, 0, stream);
}

// 2-D tensor map {feat, rows} over a row-major fp32 matrix, box = {16 columns, box_rows rows}, no swizzle
static bool make_slice_map(CUtensorMap* map, const float* base, uint64_t feat, uint64_t rows, int64_t ld,
                           uint32_t box_rows) {
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      return reinterpret_cast<EncodeTiledFn>(sym);
    return nullptr;
  }();
  if (!fn) return false;
  cuuint64_t dims[2] = {feat, rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)kWideCols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

extern "C" int cgc_gather_wide_ell(const uint16_t* ell, const float* ell_val, const float* ell_scale,
                                   const uint32_t* perm, int32_t width, const int32_t* rowptr, const int32_t* col,
                                   const float* val, const float* x, int64_t ldx, int32_t num_graphs, int32_t nmax,
                                   int32_t feat, float* out, int64_t ldo, int accumulate, void* stream) {
  CGC_REQUIRE(ell && perm && rowptr && col && x && out && num_graphs > 0 && nmax > 0 && feat > 0, CGC_ERR_ARG,
              "cgc_gather_wide_ell: bad args");
  CGC_REQUIRE(width == 8 || width == 16, CGC_ERR_UNSUPPORTED, "cgc_gather_wide_ell: width must be 8 or 16");
  CGC_REQUIRE((ell_val == nullptr) == (val == nullptr), CGC_ERR_ARG,
              "cgc_gather_wide_ell: val and ell_val must be given together");
  CGC_REQUIRE(feat % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0 && ldx >= feat && ldo >= feat && aligned16(x) &&
                  aligned16(out) && aligned16(ell) && (ell_val == nullptr || aligned16(ell_val)),
              CGC_ERR_UNSUPPORTED, "cgc_gather_wide_ell: needs feat, ldx, ldo multiples of 4 and 16B-aligned pointers");
  // box rows: as large as possible (<= 256) with the smallest overhang past nmax; two buffers if they fit
  const int64_t budget = 227 * 1024 - 256;
  int best_box = 0, nbuf = 0;
  for (int want = 2; want >= 1 && best_box == 0; --want) {
    int64_t best_rows = -1;
    for (int box = 256; box >= 8; box -= 8) {
      const int64_t rows_alloc = ceil_div(nmax, box) * box;
      const int64_t rows_buf = rows_alloc > nmax + 2 ? rows_alloc : nmax + 2;
      if (rows_buf * 64 * want > budget) continue;
      if (best_rows < 0 || rows_alloc < best_rows) { best_rows = rows_alloc; best_box = box; }
    }
    if (best_box) nbuf = want;
  }
  CGC_REQUIRE(best_box > 0, CGC_ERR_UNSUPPORTED, "cgc_gather_wide_ell: %d rows per graph do not fit shared memory", nmax);
  CUtensorMap map;
  CGC_REQUIRE(make_slice_map(&map, x, (uint64_t)feat, (uint64_t)num_graphs * nmax, ldx, (uint32_t)best_box),
              CGC_ERR_UNSUPPORTED, "cgc_gather_wide_ell: cuTensorMapEncodeTiled failed");
  const int64_t rows_alloc = ceil_div(nmax, best_box) * best_box;
  const int64_t rows_buf = rows_alloc > nmax + 2 ? rows_alloc : nmax + 2;
  const size_t smem = (size_t)(rows_buf * 64 * nbuf);
  const int64_t nitems = (int64_t)num_graphs * ceil_div(feat, kWideCols);
  const int grid = (int)(nitems < kNumSMs ? nitems : kNumSMs);
  cudaStream_t st = as_stream(stream);
#define CGC_WIDE(W_, WT_)                                                                                          \
  do {                                                                                                             \
    auto kern = k_gather_wide_ell<W_, WT_, ((W_) == 8 ? 8 : 4) / ((WT_) ? 2 : 1)>;                                                                  \
    CGC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                  \
    kern<<<grid, kWideThreads, smem, st>>>(map, ell, ell_val, ell_scale, perm, rowptr, col, val, num_graphs, nmax,  \
                                          feat, best_box, nbuf, out, ldo, accumulate);                             \
  } while (0)
  if (width == 8) { if (ell_val) CGC_WIDE(8, true); else CGC_WIDE(8, false); }
  else            { if (ell_val) CGC_WIDE(16, true); else CGC_WIDE(16, false); }
#undef CGC_WIDE
  CGC_CHECK_LAUNCH("k_gather_wide_ell");
  return 0;
}
