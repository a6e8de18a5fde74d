// tcgen05 (5th-gen tensor core) batched GEMM in 3xTF32 with fp32 "promotion" of the accumulator.
//
// fp32 operands are split in-kernel into hi (what kind::tf32 reads: the upper 19 bits, the hardware
// ignores the low 13 mantissa bits -- verified by tests/test_gpu_ops.py) and lo = a - hi, and
// D += lo*hi + hi*lo + hi*hi restores the fp32 products.  The tensor core adds into TMEM with truncation
// (round toward zero; measured: a biased ~0.5 ulp(D) per accumulation step), which over a
// K = 1000..100000 contraction would cost 1e-5..1e-3 relative.  So the accumulator is drained every
// 2 K-blocks (K = 64, 24 MMA adds): the epilogue warps read the chunk from TMEM and add it to fp32
// registers with round-to-nearest while the MMA warp fills the other TMEM buffer.  What remains is ~1e-6
// relative per chunk (plain fp32 FMA accumulation is ~1e-7; un-promoted 3xTF32 measured 1e-5..1e-3).
//
// Serves lin of GNN_Module (reference model/network.py:121-124), S^T X / S^T (A S) of _diff_pool
// (network.py:206-207), matmul(adj, x) of DenseSAGEConv on the dense stages, and all their gradients.
//
// Structure (one persistent CTA per SM, 512 threads, warp-specialised):
//   warp 0       TMA producer: cp.async.bulk.tensor of the fp32 A / B tiles into a 4-stage ring
//   warps 4-7    split: B lo tile written next to the TMA-landed B tile (element-wise, layout preserving);
//                A: each thread gathers its row out of the swizzled A tile, splits it and stores hi | lo
//                into TENSOR MEMORY (tcgen05.st).  The MMA reads A from TMEM, which removes the A operand
//                from the shared-memory read stream (shared-memory bandwidth is the measured limiter)
//   warp 1       MMA issuer: the warp runs the loop uniformly, one elected lane issues 3 x 4
//                tcgen05.mma.kind::tf32 (128x128x8, A from TMEM, B via shared-memory descriptors) per stage;
//                tcgen05.commit frees the stage and, every 2 stages, hands the accumulator to the epilogue
//   warp 2       TMEM allocation: 2 accumulator buffers x 128 columns + 4 stages x 64 operand columns = 512
//   warps 8-15   promotion + epilogue: tcgen05.ld 32x32b chunks -> fp32 registers (+=), at the end of the tile
//                alpha / bias / row-scale / row-mask / beta and 128-byte row segments to global
// K-major operands use the 128B swizzle, MN-major operands the 128B swizzle with 32B atoms (the only
// MN-major layout kind::tf32 accepts), so no transposed copy of an activation is ever made.
#include <cuda.h>
#include <stdlib.h>

#include "ell_gather.cuh"
#include "gemm.cuh"

namespace cgc {

namespace tc {

constexpr int BM = 128, BK = 32;                      // BN is a template parameter: 128, 64 or 32
// Ring depth per instantiation.  The narrow-N kernels (BN <= 64, one CTA per tile) serve the dense-stage A'x / A'^T g
// contractions, 333 MB streaming reads with 4 - 8 KB of B per stage: with 4 stages every stage spent ~3 700 cycles
// between its TMA issue and its release (DRAM flight + split + MMA) and the ring, not DRAM or the split warps
// (60 % busy) or the MMA warp (59 %), set the pace: 1 200 cycles per stage, 0.55 - 0.64 of the HBM peak
// (profiles/gemm_dense_r2_ncu.txt).  Their accumulators need only 2 x BN <= 128 tensor-memory columns, which leaves
// room for 6 staged A tiles (128 + 6 x 64 = 512) and 6 x 32 KB of shared memory.
__host__ __device__ constexpr int stages_of(int bn, bool pair) { return (!pair && bn <= 64) ? 6 : 4; }
__host__ __device__ constexpr int a_tmem_base_of(int bn, bool pair) { return (!pair && bn <= 64) ? 128 : 256; }
constexpr int kOperandBytes = BM * BK * 4;            // A tile: 16 KB.  Stage = A (as landed) | B (as landed = B_hi) | B_lo
constexpr int kThreads = 640;                          // warps 0-3 control, 4-7 split group 0, 8-15 promotion, 16-19 split group 1
constexpr int kSplitThreads = 128;                    // threads of ONE split group (= rows of the A tile)
constexpr int kSplitGroups = 2;                       // the groups take alternate stages of the ring
constexpr int kEpiThreads = 256;
constexpr int kDefaultChunkBlocks = 2;               // K-blocks accumulated in TMEM before promotion to fp32 registers
                                                      // tensor memory: accumulators (2 buffers x BN columns) in
                                                      // [0, a_tmem_base); then the A operand tiles, per stage hi[32] | lo[32]
constexpr int kATmemCols = 2 * BK;
constexpr int kTmemCols = 512;                        // 2 * 128 accumulator + 4 * 64 operand columns
constexpr size_t smem_bytes(int bn, bool pair) {
  return (size_t)stages_of(bn, pair) * (kOperandBytes + 2 * (pair ? bn / 2 : bn) * BK * 4) + 1024 /*align*/ + 256 /*barriers*/;
}

struct Params {
  GemmParams g;
  int a_mn_major, b_mn_major;
  int tiles_m, tiles_n, total_tiles;       // total = tiles_m * tiles_n * batch * ksplit
  int c_vec;
  int passes;        // 3: 3xTF32 (lo*hi + hi*lo + hi*hi, fp32-grade products); 1: single TF32 pass on operands
                     //    rounded to nearest (10-bit mantissa products, unbiased)
  int chunk_blocks;  // K-blocks per TMEM accumulation chunk (promotion period): 24 MMA adds per chunk either way
  int bn;            // N tile (= the kernel's BN template argument)
  int b_warps;       // pair mode: B lo split on warps 2-3 (lab switch CGC_TC_BWARPS, CGC_LAB builds only)
  int a_first;       // split warps stage A before splitting B (lab switch CGC_TC_AFIRST, CGC_LAB builds only)
};

// fp32 -> tf32 (10-bit mantissa), round to nearest (ties away from zero): what a single-pass contraction feeds the
// tensor core instead of the hardware's truncation, so that sign-definite operands (softmax rows) carry no bias
__device__ __forceinline__ uint32_t tf32_rn(uint32_t x) { return (x + 0x1000u) & 0xFFFFE000u; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "TC_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra TC_DONE;\n\t"
      "bra TC_WAIT;\n\t"
      "TC_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// ---- CTA-pair (cta_group::2) helpers
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster.  Default (CTA-scope
// release) semantics on purpose: `.release.cluster` / `.acquire.cluster` compile to MEMBAR.ALL.GPU + ERRBAR on the
// arrive and CCTL.IVALL on every wait (measured: the pair kernel 1.5x SLOWER than the single-CTA one); the
// tensor-memory / shared-memory hand-off is ordered by the tcgen05 fences and fence.proxy.async around it.
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(rank)
      : "memory");
}
// wait on a barrier whose arrivals may come from the peer CTA (same instruction as mbar_wait, own labels)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "TCC_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra TCC_DONE;\n\t"
      "bra TCC_WAIT;\n\t"
      "TCC_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}

// UMMA shared-memory descriptor (sm_100 version bit set).  layout_type: 2 = SWIZZLE_128B (K-major),
// 1 = SWIZZLE_128B_BASE32B (MN-major tf32).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;          // descriptor version (Blackwell)
  d |= (uint64_t)layout_type << 61;
  return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand from tensor memory (lane = row of the tile, column = K index), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same, with the B descriptor passed as two 32-bit halves (no 64-bit arithmetic in the issue loop)
__device__ __forceinline__ void umma_tf32_ts2(uint32_t tmem_d, uint32_t tmem_a, uint32_t bdesc_lo, uint32_t bdesc_hi,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 bd;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 bd, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %4, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "r"(bdesc_lo), "r"(bdesc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// CTA pair: D (256 x N, 128 rows in each CTA's TMEM) += A (each CTA's TMEM) * B (N/2 columns in each CTA's smem)
__device__ __forceinline__ void umma_tf32_ts2_pair(uint32_t tmem_d, uint32_t tmem_a, uint32_t bdesc_lo,
                                                   uint32_t bdesc_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 bd;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 bd, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], bd, %4, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "r"(bdesc_lo), "r"(bdesc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}
// 32 consecutive TMEM columns of this thread's lane <- 32 registers
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// commit of cta_group::2 MMAs, arriving on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// acc[0..15] += TMEM[lane, col .. col+15]
__device__ __forceinline__ void tmem_ld16_add(uint32_t taddr, float* acc) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] += __uint_as_float(r[i]);
}
// acc[0..31] += TMEM[lane, col .. col+31]
__device__ __forceinline__ void tmem_ld32_add(uint32_t taddr, float* acc) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] += __uint_as_float(r[i]);
}

struct TileCoord {
  int bz, split, m0, n0, kbeg, kend;
};
// bm = rows of one scheduled tile (128, or 256 for a CTA pair); rank = this CTA's half of it
__device__ __forceinline__ TileCoord tile_coord(const Params& p, int t, int bm = BM, int rank = 0) {
  TileCoord c;
  int nt = t % p.tiles_n;
  int r = t / p.tiles_n;
  int mt = r % p.tiles_m;
  int z = r / p.tiles_m;                  // bz * ksplit + split
  c.bz = z / p.g.ksplit;
  c.split = z - c.bz * p.g.ksplit;
  c.m0 = mt * bm + rank * BM;
  c.n0 = nt * p.bn;
  c.kbeg = c.split * p.g.kchunk;
  c.kend = min(p.g.k, c.kbeg + p.g.kchunk);
  return c;
}

// final write of NC (16 or 32) consecutive columns of one output row (v already holds the fp32 sums)
template <int NC>
__device__ __forceinline__ void store_segment(const Params& p, const TileCoord& c, int row, int col0, const float* v) {
  const GemmParams& g = p.g;
  if (row >= g.m || col0 >= g.n) return;
  if (g.ksplit > 1) {
    float* dst = g.partial + (((int64_t)(c.bz * g.ksplit + c.split)) * g.m + row) * g.n + col0;
#pragma unroll
    for (int j = 0; j < NC; ++j)
      if (col0 + j < g.n) dst[j] = g.alpha * v[j];
    return;
  }
  const int64_t grow = (int64_t)c.bz * g.m + row;
  const float rsc = g.row_scale ? __ldg(g.row_scale + grow) : 1.0f;
  const bool valid = g.num_nodes ? row_valid(g.num_nodes, g.nmax, grow) : true;
  float* dst = g.c + (int64_t)c.bz * g.stride_c + (int64_t)row * g.ldc + col0;
  if (p.c_vec && col0 + NC <= g.n) {
#pragma unroll
    for (int j = 0; j < NC; j += 4) {
      float4 o;
      float* ov = reinterpret_cast<float*>(&o);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float x = g.alpha * v[j + u];
        if (g.bias) x += __ldg(g.bias + col0 + j + u);
        x *= rsc;
        ov[u] = valid ? x : 0.f;
      }
      if (g.beta != 0.f) {
        const float4 old = *reinterpret_cast<const float4*>(dst + j);
        o.x += g.beta * old.x; o.y += g.beta * old.y; o.z += g.beta * old.z; o.w += g.beta * old.w;
      }
      *reinterpret_cast<float4*>(dst + j) = o;
    }
  } else {
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      if (col0 + j < g.n) {
        float x = g.alpha * v[j];
        if (g.bias) x += __ldg(g.bias + col0 + j);
        x *= rsc;
        if (!valid) x = 0.f;
        if (g.beta != 0.f) x += g.beta * dst[j];
        dst[j] = x;
      }
    }
  }
}

// PAIR: two CTAs of a cluster (one TPC) compute a 256 x BN tile with tcgen05.mma.cta_group::2: each CTA stages
// its own 128 rows of A (TMEM) and only BN/2 columns of B (shared memory); the leader (rank 0) issues the MMAs,
// whose B reads are served half by each CTA's shared memory -- the B share of the shared-memory traffic halves.
template <int BN, bool PAIR>
__global__ void __launch_bounds__(kThreads, 1)
k_gemm_tc(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const Params p) {
  constexpr int kStages = stages_of(BN, PAIR);
  constexpr int kATmemBase = a_tmem_base_of(BN, PAIR);
  constexpr int kBCols = PAIR ? BN / 2 : BN;          // B columns staged by this CTA
  constexpr int kBBytes = kBCols * BK * 4;
  constexpr int kTileM = PAIR ? 2 * BM : BM;
  constexpr int kStageBytes = kOperandBytes + 2 * kBBytes;
  constexpr int kTxBytes = kOperandBytes + kBBytes;
  constexpr int kAccCols = BN;
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte alignment by offset (keeps the pointer in the shared address space: LDS/STS, not generic LD/ST)
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * kStageBytes);
  uint64_t* full = bars;                  // TMA -> split warps
  uint64_t* xfull = bars + kStages;       // split warps -> MMA
  uint64_t* empty = bars + 2 * kStages;   // MMA -> TMA
  uint64_t* tfull = bars + 3 * kStages;   // MMA -> promotion warps (2 accumulator buffers)
  uint64_t* tempty = tfull + 2;           // promotion warps -> MMA (2)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = PAIR ? (int)cluster_ctarank() : 0;
  const int tile_first = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&xfull[s], PAIR ? 2 * (kSplitThreads / 32 + (p.b_warps ? 2 : 0)) : kSplitThreads);   // pair: one arrival per A / B split warp of both CTAs
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], PAIR ? 2 * (kEpiThreads / 32) : kEpiThreads);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == 2) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"((uint32_t)kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"((uint32_t)kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();            // the peer's barriers exist before anything arrives on them remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = tile_first; t < p.total_tiles; t += tile_step) {
        const TileCoord c = tile_coord(p, t, kTileM, rank);
        const int bn0 = c.n0 + rank * kBCols;          // pair: this CTA's half of the B columns
        for (int k0 = c.kbeg; k0 < c.kend; k0 += BK) {
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], kTxBytes);
          unsigned char* sa = smem + (size_t)s * kStageBytes;
          unsigned char* sb = sa + kOperandBytes;
          if (p.a_mn_major) {
#pragma unroll
            for (int ch = 0; ch < BM / 32; ++ch) tma_load_3d(sa + ch * 4096, &map_a, c.m0 + ch * 32, k0, c.bz, &full[s]);
          } else {
            tma_load_3d(sa, &map_a, k0, c.m0, c.bz, &full[s]);
          }
          if (p.b_mn_major) {
#pragma unroll
            for (int ch = 0; ch < kBCols / 32; ++ch) tma_load_3d(sb + ch * 4096, &map_b, bn0 + ch * 32, k0, c.bz, &full[s]);
          } else {
            tma_load_3d(sb, &map_b, k0, bn0, c.bz, &full[s]);
          }
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    // The whole warp runs the loop (so addresses / descriptors stay warp-uniform and cheap to form);
    // one elected lane issues.  A is fed from TMEM (lane = row, column = k), B from shared memory.
    // Pair: only the leader CTA issues (for both); its waits see the peer's arrivals (cluster scope).
    if (!PAIR || rank == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.b_mn_major << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
      // B MN-major (SW128 with 32B atoms): 32-element chunks 4096 B apart (LBO), 4-k-row atoms 512 B apart
      //           (SBO), one K=8 step = the next 1024 B.
      // B K-major  (SW128): 8-row groups 1024 B apart (SBO), one K=8 step = 32 B inside the 128 B swizzle row.
      const uint32_t b_lbo = p.b_mn_major ? 4096u : 16u;
      const uint32_t b_sbo = p.b_mn_major ? 512u : 1024u;
      const uint32_t b_lt = p.b_mn_major ? 1u : 2u;
      const uint32_t b_kstep16 = (p.b_mn_major ? 1024u : 32u) >> 4;       // K step in descriptor address units
      // descriptor = {lo: addr>>4 | lbo>>4 << 16, hi: sbo>>4 | version 1 << 14 | layout << 29}
      const uint32_t desc_hi = (b_sbo >> 4) | (1u << 14) | (b_lt << 29);
      const uint32_t desc_lo_base = (b_lbo >> 4) << 16;
      const uint32_t smem_base16 = smem_u32(smem) >> 4;
      int s = 0;
      uint32_t ph = 0;
      uint32_t chunk_it = 0;
      for (int t = tile_first; t < p.total_tiles; t += tile_step) {
        const TileCoord c = tile_coord(p, t, kTileM, rank);
        int k0 = c.kbeg;
        while (k0 < c.kend) {
          const uint32_t buf = chunk_it & 1u;
          if (PAIR) mbar_wait_cluster(&tempty[buf], ((chunk_it >> 1) & 1u) ^ 1u);
          else mbar_wait(&tempty[buf], ((chunk_it >> 1) & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * (uint32_t)kAccCols;
          uint32_t accumulate = 0;
          for (int kb = 0; kb < p.chunk_blocks && k0 < c.kend; ++kb, k0 += BK) {
            if (PAIR) mbar_wait_cluster(&xfull[s], ph);
            else mbar_wait(&xfull[s], ph);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t bh = desc_lo_base | (smem_base16 + (uint32_t)((s * kStageBytes + kOperandBytes) >> 4));
              const uint32_t bl = bh + (uint32_t)(kBBytes >> 4);
              const uint32_t a_hi = tmem_base + (uint32_t)(kATmemBase + s * kATmemCols);
              const uint32_t a_lo = a_hi + (uint32_t)BK;
#pragma unroll
              if (p.passes == 1) {
#pragma unroll
                for (int kk = 0; kk < BK / 8; ++kk) {
                  if (PAIR) umma_tf32_ts2_pair(d_tmem, a_hi + kk * 8, bh + kk * b_kstep16, desc_hi, idesc, kk == 0 ? accumulate : 1u);
                  else umma_tf32_ts2(d_tmem, a_hi + kk * 8, bh + kk * b_kstep16, desc_hi, idesc, kk == 0 ? accumulate : 1u);
                }
              } else {
#pragma unroll
              for (int kk = 0; kk < BK / 8; ++kk) {
                if (PAIR) {
                  umma_tf32_ts2_pair(d_tmem, a_lo + kk * 8, bh + kk * b_kstep16, desc_hi, idesc, kk == 0 ? accumulate : 1u);
                  umma_tf32_ts2_pair(d_tmem, a_hi + kk * 8, bl + kk * b_kstep16, desc_hi, idesc, 1u);
                  umma_tf32_ts2_pair(d_tmem, a_hi + kk * 8, bh + kk * b_kstep16, desc_hi, idesc, 1u);
                } else {
                  umma_tf32_ts2(d_tmem, a_lo + kk * 8, bh + kk * b_kstep16, desc_hi, idesc, kk == 0 ? accumulate : 1u);
                  umma_tf32_ts2(d_tmem, a_hi + kk * 8, bl + kk * b_kstep16, desc_hi, idesc, 1u);
                  umma_tf32_ts2(d_tmem, a_hi + kk * 8, bh + kk * b_kstep16, desc_hi, idesc, 1u);
                }
              }
              }
              if (PAIR) umma_commit_pair(&empty[s]);    // frees the stage in BOTH CTAs
              else umma_commit(&empty[s]);              // frees the stage once these MMAs have read it
            }
            __syncwarp();
            accumulate = 1u;
            if (++s == kStages) { s = 0; ph ^= 1; }
          }
          if (elect_one()) {                            // chunk complete -> promotion warps (pair: of both CTAs)
            if (PAIR) umma_commit_pair(&tfull[buf]);
            else umma_commit(&tfull[buf]);
          }
          __syncwarp();
          ++chunk_it;
        }
      }
    }
  } else if ((warp >= 4 && warp < 8) || warp >= 16) {
    // ================================ hi / lo split ================================
    // Two groups of four warps (4-7 and 16-19; a warp reaches the TMEM lane quarter warp % 4, so each group covers
    // the 128 rows) take ALTERNATE stages: one group alone spent ~800 cycles per stage in this dependent chain
    // (LDS of its row -> split -> tcgen05.st -> wait::st -> fences -> arrive) and was busy 86 % of the time while
    // the tensor pipe idled -- the contraction ran at the split rate whatever the number of MMA passes.
    const int group = warp >= 16 ? 1 : 0;
    const int tid = threadIdx.x - (group ? 512 : 128);
    int s = 0;
    uint32_t ph = 0;
    uint32_t stage_it = 0;                              // stages seen by this warp, over all its tiles
    for (int t = tile_first; t < p.total_tiles; t += tile_step) {
      const TileCoord c = tile_coord(p, t, kTileM, rank);
      for (int k0 = c.kbeg; k0 < c.kend; k0 += BK, ++stage_it) {
        if ((int)(stage_it % kSplitGroups) != group) {  // the other group's stage: just keep the ring position
          if (++s == kStages) { s = 0; ph ^= 1; }
          continue;
        }
        mbar_wait(&full[s], ph);
        unsigned char* sa = smem + (size_t)s * kStageBytes;
        auto split_b = [&]() {
        // ---- B: lo tile next to the landed tile (element-wise, layout preserving); pair: done by warps 2-3
        uint4* bhi = reinterpret_cast<uint4*>(sa + kOperandBytes);
        uint4* blo = bhi + kBBytes / 16;
#pragma unroll 4
        for (int i = (PAIR && p.b_warps) ? kBBytes / 16 : tid; i < kBBytes / 16; i += kSplitThreads) {
          const uint4 v = bhi[i];
          if (p.passes == 1) {                         // single pass: the landed tile is rounded in place, no lo tile
            uint4 h;
            h.x = tf32_rn(v.x); h.y = tf32_rn(v.y); h.z = tf32_rn(v.z); h.w = tf32_rn(v.w);
            bhi[i] = h;
            continue;
          }
          uint4 l;
          l.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(v.x & 0xFFFFE000u));
          l.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(v.y & 0xFFFFE000u));
          l.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(v.z & 0xFFFFE000u));
          l.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(v.w & 0xFFFFE000u));
          blo[i] = l;
        }
        };
        if (!p.a_first) split_b();
        // ---- A: thread `tid` owns row `tid` of the tile: gather its 32 k-values out of the swizzled
        //      shared-memory tile, split, and store hi | lo into this lane's TMEM columns.
        uint32_t ah[BK], al[BK];
        if (p.a_mn_major) {
          // box (tid / 32) of 4096 B; k-row k at k * 128 B; 32-byte units XOR-swizzled with (k & 3)
          const unsigned char* base = sa + (tid >> 5) * 4096 + (tid & 7) * 4;
          const int unit = (tid & 31) >> 3;
#pragma unroll
          for (int k = 0; k < BK; ++k)
            ah[k] = *reinterpret_cast<const uint32_t*>(base + k * 128 + ((unit ^ (k & 3)) << 5));
        } else {
          // row tid at tid * 128 B; 16-byte chunks XOR-swizzled with (tid & 7)
          const uint4* rowp = reinterpret_cast<const uint4*>(sa + tid * 128);
#pragma unroll
          for (int ch = 0; ch < BK / 4; ++ch) {
            const uint4 v = rowp[ch ^ (tid & 7)];
            ah[4 * ch + 0] = v.x; ah[4 * ch + 1] = v.y; ah[4 * ch + 2] = v.z; ah[4 * ch + 3] = v.w;
          }
        }
        const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(kATmemBase + s * kATmemCols);
        if (p.passes == 1) {
#pragma unroll
          for (int k = 0; k < BK; ++k) ah[k] = tf32_rn(ah[k]);
          tmem_st32(ta, ah);
        } else {
#pragma unroll
          for (int k = 0; k < BK; ++k) {
            const uint32_t h = ah[k] & 0xFFFFE000u;
            al[k] = __float_as_uint(__uint_as_float(ah[k]) - __uint_as_float(h));
            ah[k] = h;
          }
          tmem_st32(ta, ah);
          tmem_st32(ta + (uint32_t)BK, al);
        }
        if (p.a_first) split_b();                      // overlaps the latency of the two tensor-memory stores just issued
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (PAIR) {
          // one arrival per warp on the LEADER's barrier (remote arrivals are DSMEM transactions: 256 of them
          // per stage would serialise on the barrier)
          __syncwarp();
          if (lane == 0) {
            if (rank == 0) mbar_arrive(&xfull[s]);
            else mbar_arrive_cluster(&xfull[s], 0);
          }
        } else {
          mbar_arrive(&xfull[s]);
        }
        if (++s == kStages) { s = 0; ph ^= 1; }
      }
    }
  } else if (PAIR && p.b_warps && (warp == 2 || warp == 3)) {
    // ================================ B lo split (pair mode) ================================
    // 64 otherwise idle threads take the element-wise B split off the A-staging warps (measured: those were busy
    // 87 % of the time and paced the MMA).
    const int tid = threadIdx.x - 64;
    int s = 0;
    uint32_t ph = 0;
    for (int t = tile_first; t < p.total_tiles; t += tile_step) {
      const TileCoord c = tile_coord(p, t, kTileM, rank);
      for (int k0 = c.kbeg; k0 < c.kend; k0 += BK) {
        mbar_wait(&full[s], ph);
        uint4* bhi = reinterpret_cast<uint4*>(smem + (size_t)s * kStageBytes + kOperandBytes);
        uint4* blo = bhi + kBBytes / 16;
#pragma unroll
        for (int i = tid; i < kBBytes / 16; i += 64) {
          const uint4 v = bhi[i];
          if (p.passes == 1) {
            uint4 h;
            h.x = tf32_rn(v.x); h.y = tf32_rn(v.y); h.z = tf32_rn(v.z); h.w = tf32_rn(v.w);
            bhi[i] = h;
            continue;
          }
          uint4 l;
          l.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(v.x & 0xFFFFE000u));
          l.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(v.y & 0xFFFFE000u));
          l.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(v.z & 0xFFFFE000u));
          l.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(v.w & 0xFFFFE000u));
          blo[i] = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          if (rank == 0) mbar_arrive(&xfull[s]);
          else mbar_arrive_cluster(&xfull[s], 0);
        }
        if (++s == kStages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp >= 8 && warp < 16) {
    // ================================ promotion + epilogue ================================
    const int q = (warp - 8) & 3;                      // TMEM lane quarter == warp_id % 4
    const int half = (warp - 8) >> 2;                  // which BN/2 accumulator columns this warp owns
    constexpr int kCols = BN / 2;                      // 64, 32 or 16
    uint32_t chunk_it = 0;
    for (int t = tile_first; t < p.total_tiles; t += tile_step) {
      const TileCoord c = tile_coord(p, t, kTileM, rank);
      float acc[kCols];
#pragma unroll
      for (int j = 0; j < kCols; ++j) acc[j] = 0.f;
      for (int k0 = c.kbeg; k0 < c.kend; k0 += p.chunk_blocks * BK) {
        const uint32_t buf = chunk_it & 1u;
        mbar_wait(&tfull[buf], (chunk_it >> 1) & 1u);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)kAccCols + (uint32_t)(half * kCols);
        // 16 columns per tcgen05.ld: the 64 accumulators plus a 32-register landing zone would not fit the 96
        // registers a 640-thread CTA leaves each thread
#pragma unroll
        for (int j = 0; j < kCols; j += 16) tmem_ld16_add(taddr + (uint32_t)j, acc + j);
        tc_fence_before();
        if (PAIR) {
          __syncwarp();
          if (lane == 0) {
            if (rank == 0) mbar_arrive(&tempty[buf]);
            else mbar_arrive_cluster(&tempty[buf], 0);
          }
        } else {
          mbar_arrive(&tempty[buf]);
        }
        ++chunk_it;
      }
      const int row = c.m0 + q * 32 + lane;
      if constexpr (kCols == 16) {
        store_segment<16>(p, c, row, c.n0 + half * kCols, acc);
      } else {
#pragma unroll
        for (int j = 0; j < kCols; j += 32) store_segment<32>(p, c, row, c.n0 + half * kCols + j, acc + j);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();            // no CTA frees tensor memory (or exits) while its peer still uses the pair
  if (warp == 2) {
    tc_fence_after();
    if (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemCols)
                   : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemCols)
                   : "memory");
  }
}

__global__ void __launch_bounds__(256) k_tc_splitk_reduce(GemmParams p) {
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

// ------------------------------------------------------------------------------------ host side
// 3-D tensor map {inner, rows, batch} over a row-major fp32 operand (encoded once per distinct view, api.cu)
static bool make_map(CUtensorMap* map, const float* base, uint64_t inner, uint64_t rows, uint64_t batch, int64_t ld,
                     int64_t batch_stride, uint32_t box_rows, bool mn_major) {
  return tensor_map_3d(map, base, inner, rows, batch, ld, batch_stride, box_rows, mn_major);
}

}  // namespace tc

bool gemm_tc_supported(const GemmParams& p, int ta, int tb, bool forced) {
  (void)ta;
  (void)tb;
  if (!p.a_vec || !p.b_vec) return false;                      // 16-byte aligned base, ld and batch stride
  if (!forced) {                                               // small problems stay on the FMA path
    // skinny outputs (n or m of 16..95) are taken when K is long: those contractions are bound by
    // streaming the big operand, which the TMA pipeline does well even with a mostly idle MMA tile
    const int lo = p.m < p.n ? p.m : p.n, hi = p.m < p.n ? p.n : p.m;
    if (hi < 96 || lo < 16 || p.k < 32) return false;
    if (lo < 96 && p.k < 512) return false;
    if ((double)hi * (lo < 96 ? 128 : lo) * p.k * p.batch < 6.7e7) return false;
  }
  if (p.lda * 4 >= (1ll << 40) || p.ldb * 4 >= (1ll << 40)) return false;
  return tensor_map_available();
}

int gemm_tc(GemmParams p, int ta, int tb, int passes, void* work, int64_t work_bytes, cudaStream_t st) {
  using namespace tc;
  // promotion period: 24 tensor-core adds per TMEM chunk in both modes (2 K-blocks x 12 MMAs, or 6 x 4)
  static const int chunk_override = lab_env("CGC_TC_CHUNK", 0);
  int chunk_blocks = passes == 1 ? 3 * kDefaultChunkBlocks : kDefaultChunkBlocks;
  if (chunk_override > 0) chunk_blocks = chunk_override > 64 ? 64 : chunk_override;
  Params q;
  q.passes = passes == 1 ? 1 : 3;
  q.chunk_blocks = chunk_blocks;
  q.a_mn_major = ta ? 1 : 0;          // A stored [k, m]: m contiguous
  q.b_mn_major = tb ? 0 : 1;          // B stored [k, n]: n contiguous
  const int bn = p.n <= 32 ? 32 : (p.n <= 64 ? 64 : 128);      // skinny outputs: narrow MMA tile, same A stream
  q.bn = bn;
  static const int pair_mode = lab_env("CGC_TC_PAIR", 1);        // 0 disables the cta_group::2 kernel (CGC_LAB builds)
  const bool pair = pair_mode != 0 && bn == 128 && p.m > BM;    // 256-row tiles on a CTA pair
  q.tiles_m = (int)ceil_div(p.m, pair ? 2 * BM : BM);
  q.tiles_n = (int)ceil_div(p.n, bn);
  int64_t tiles = (int64_t)q.tiles_m * q.tiles_n * p.batch;
  p.ksplit = 1;
  p.kchunk = p.k;
  p.partial = nullptr;
  const int units = pair ? kNumSMs / 2 : kNumSMs;               // CTAs (or CTA pairs) that run tiles concurrently
  if (tiles < units && p.k >= 4096 && work != nullptr) {        // split K to fill the SMs
    // pick the split whose work items fill whole waves best (items / units close to an integer from below)
    int64_t kmax = p.k / 1024;
    if (kmax > 24) kmax = 24;
    int64_t ks = 1;
    double best = 0.0;
    for (int64_t cand = 1; cand <= kmax; ++cand) {
      if (cand * p.batch * (int64_t)p.m * p.n * 4 > work_bytes) break;
      const double waves = (double)(tiles * cand) / units;
      const double eff = waves / (double)ceil_div(tiles * cand, units);
      if (eff > best + 0.02) { best = eff; ks = cand; }
    }
    if (ks > 1) {
      int64_t chunk = ceil_div(ceil_div(p.k, ks), chunk_blocks * BK) * (chunk_blocks * BK);
      p.kchunk = (int)chunk;
      p.ksplit = (int)ceil_div(p.k, chunk);
      p.partial = reinterpret_cast<float*>(work);
    }
  }
  tiles *= p.ksplit;
  CGC_REQUIRE(tiles < (1ll << 31), CGC_ERR_UNSUPPORTED, "cgc_gemm(tc): too many tiles");
  q.total_tiles = (int)tiles;
  q.c_vec = (aligned16(p.c) && p.ldc % 4 == 0 && p.stride_c % 4 == 0) ? 1 : 0;
  static const int b_warps = lab_env("CGC_TC_BWARPS", 0);
  static const int a_first = lab_env("CGC_TC_AFIRST", 0);
  q.b_warps = b_warps;
  q.a_first = a_first;
  q.g = p;

  CUtensorMap map_a, map_b;
  bool ok;
  if (q.a_mn_major) ok = make_map(&map_a, p.a, (uint64_t)p.m, (uint64_t)p.k, (uint64_t)p.batch, p.lda, p.stride_a, BK, true);
  else ok = make_map(&map_a, p.a, (uint64_t)p.k, (uint64_t)p.m, (uint64_t)p.batch, p.lda, p.stride_a, BM, false);
  CGC_REQUIRE(ok, CGC_ERR_UNSUPPORTED, "cgc_gemm(tc): cuTensorMapEncodeTiled failed for A");
  if (q.b_mn_major) ok = make_map(&map_b, p.b, (uint64_t)p.n, (uint64_t)p.k, (uint64_t)p.batch, p.ldb, p.stride_b, BK, true);
  else ok = make_map(&map_b, p.b, (uint64_t)p.k, (uint64_t)p.n, (uint64_t)p.batch, p.ldb, p.stride_b,
                     (uint32_t)(pair ? bn / 2 : bn), false);
  CGC_REQUIRE(ok, CGC_ERR_UNSUPPORTED, "cgc_gemm(tc): cuTensorMapEncodeTiled failed for B");

  int grid = q.total_tiles < kNumSMs ? q.total_tiles : kNumSMs;
  const size_t smem = smem_bytes(bn, pair);
  if (pair) {
    // one cluster of 2 CTAs (same TPC) per 256-row tile; persistent over the tiles
    const int clusters = q.total_tiles < kNumSMs / 2 ? q.total_tiles : kNumSMs / 2;
    auto kern = k_gemm_tc<128, true>;
    static size_t smem_set[kMaxDevices] = {};           // the size depends on the instantiation only: set once per device
    CGC_CUDA(ensure_dynamic_smem(kern, smem, smem_set));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * clusters));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CGC_CUDA(cudaLaunchKernelEx(&cfg, kern, map_a, map_b, q));
    CGC_CHECK_LAUNCH("k_gemm_tc<pair>");
  } else {
#define CGC_TC_CASE(N)                                                                                          \
  case N: {                                                                                                     \
    static size_t smem_set[kMaxDevices] = {};                                                                   \
    CGC_CUDA(ensure_dynamic_smem(k_gemm_tc<N, false>, smem, smem_set));                                         \
    k_gemm_tc<N, false><<<grid, kThreads, smem, st>>>(map_a, map_b, q);                                         \
  } break;
    switch (bn) {
      CGC_TC_CASE(32)
      CGC_TC_CASE(64)
      CGC_TC_CASE(128)
    }
#undef CGC_TC_CASE
    CGC_CHECK_LAUNCH("k_gemm_tc");
  }
  if (p.ksplit > 1) {
    int64_t total = (int64_t)p.m * p.n * p.batch;
    int64_t blocks = ceil_div(total, 256);
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    k_tc_splitk_reduce<<<(unsigned)blocks, 256, 0, st>>>(p);
    CGC_CHECK_LAUNCH("k_tc_splitk_reduce");
  }
  return 0;
}

}  // namespace cgc
