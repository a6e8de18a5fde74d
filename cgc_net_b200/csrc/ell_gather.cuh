// Device helpers shared by the shared-memory gathers over the ELL16 table (spmm.cu, sage_layer.cu):
// mbarrier / bulk-copy (TMA) wrappers, packed f32x2 arithmetic, programmatic dependent launch, and the host-side
// tensor-map cache.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace cgc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D bulk async copy global -> shared, completion counted in bytes on an mbarrier (TMA engine)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// The same copy delivered to the same shared-memory offset of every CTA of the cluster named in `mask` (one L2 read
// for all of them); each destination's mbarrier at the offset of `bar` counts the bytes
__device__ __forceinline__ void bulk_g2s_multicast(void* dst, const void* src, uint32_t bytes, uint64_t* bar,
                                                   uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 2-D tiled TMA load (box described by the tensor map) global -> shared
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

// Programmatic dependent launch (PDL).  A kernel launched with the programmatic-stream-serialization attribute may
// start while its predecessor in the stream is still running; everything it does before pdl_wait() must not touch
// memory the predecessor writes (or reads memory this kernel writes).  pdl_wait() returns once the predecessor has
// completed and its writes are visible.  pdl_launch_dependents() lets the NEXT such kernel start its prologue.
// Both are no-ops for a normally launched kernel.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ uint4 ldg_u4(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ float4 ldg_f4(const void* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// 16 bytes of shared memory as two packed f32x2 registers (operands of FADD2 / FFMA2)
struct F32x4 {
  unsigned long long lo, hi;
};
__device__ __forceinline__ F32x4 lds_x4(uint32_t addr) {
  F32x4 v;
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(v.lo), "=l"(v.hi) : "r"(addr));
  return v;
}
__device__ __forceinline__ void add_x4(F32x4& acc, const F32x4& v) {
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(acc.lo) : "l"(v.lo));
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(acc.hi) : "l"(v.hi));
}
__device__ __forceinline__ void fma_x4(F32x4& acc, float w, const F32x4& v) {
  unsigned long long ww;
  asm("mov.b64 %0, {%1, %1};" : "=l"(ww) : "f"(w));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc.lo) : "l"(v.lo), "l"(ww));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc.hi) : "l"(v.hi), "l"(ww));
}
__device__ __forceinline__ float4 unpack_x4(const F32x4& v) {
  float4 o;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(o.x), "=f"(o.y) : "l"(v.lo));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(o.z), "=f"(o.w) : "l"(v.hi));
  return o;
}

// ELL16 table: 12-bit local column ids; slot 0 carries the served local row (3 bits) above them
constexpr uint32_t kEllColMask = 0x0FFFu;
constexpr int kEllRowShift = 12;
// slot t (compile-time after unrolling) of a row's packed uint16 index words
__device__ __forceinline__ uint32_t ell_slot(const uint32_t* words, int t) {
  const uint32_t w = words[t >> 1];
  if (t == 0) return w & kEllColMask;
  return (t & 1) ? (w >> 16) : (w & 0xFFFFu);
}
// local row (0..7 within the 8-row block) served by the table row whose first index word is w0
__device__ __forceinline__ int ell_served_row(uint32_t w0) { return (int)((w0 >> kEllRowShift) & 7u); }

// ---- host side
// 2-D tensor map {cols, rows} over a row-major fp32 matrix with leading dimension ld (floats), box = {box_cols,
// box_rows}, no swizzle.  Maps are cached per thread (key = every argument), so a training loop encodes each
// distinct operand view once.  Returns false when cuTensorMapEncodeTiled is unavailable or rejects the shape.
bool tensor_map_2d(CUtensorMap* map, const float* base, uint64_t cols, uint64_t rows, int64_t ld, uint32_t box_cols,
                   uint32_t box_rows);
// 3-D map {inner, rows, batch} of the tcgen05 GEMM operands: box {32, box_rows, 1}, 128B swizzle (32B atoms when
// the operand is MN-major)
bool tensor_map_3d(CUtensorMap* map, const float* base, uint64_t inner, uint64_t rows, uint64_t batch, int64_t ld,
                   int64_t batch_stride, uint32_t box_rows, bool mn_major);
bool tensor_map_available();

// Box rows for staging `nmax` rows of `row_bytes` each through 2-D TMA boxes (<= 256 rows, multiple of 8): the
// largest box with the smallest overhang past nmax whose `nbuf` buffers fit `budget` bytes.  Returns 0 if none fits;
// *rows_buf = rows of one buffer (boxes * box_rows, at least nmax + 2 for the two zero rows).
inline int pick_box_rows(int nmax, int row_bytes, int nbuf, int64_t budget, int64_t* rows_buf) {
  int best_box = 0;
  int64_t best_rows = -1;
  for (int box = 256; box >= 8; box -= 8) {
    const int64_t rows_alloc = ceil_div(nmax, box) * box;
    const int64_t rb = rows_alloc > nmax + 2 ? rows_alloc : nmax + 2;
    if (rb * row_bytes * nbuf > budget) continue;
    if ((int64_t)box * row_bytes % 128 != 0) continue;          // every box lands 128-byte aligned
    if (best_rows < 0 || rows_alloc < best_rows) { best_rows = rows_alloc; best_box = box; *rows_buf = rb; }
  }
  return best_box;
}

// Launch attribute list for a kernel that uses pdl_wait(): programmatic stream serialization.
struct PdlLaunch {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[2];
  // cluster > 1: thread-block clusters of that many consecutive CTAs
  PdlLaunch(dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, int cluster = 1) {
    cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    int n = 0;
    if (pdl) {
      attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[n].val.programmaticStreamSerializationAllowed = 1;
      ++n;
    }
    if (cluster > 1) {
      attr[n].id = cudaLaunchAttributeClusterDimension;
      attr[n].val.clusterDim.x = (unsigned)cluster;
      attr[n].val.clusterDim.y = 1;
      attr[n].val.clusterDim.z = 1;
      ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
  }
};

}  // namespace cgc
