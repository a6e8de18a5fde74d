// Library-level entry points: version, error string, workspace size, and the cgc_gemm dispatcher.
#include <stdarg.h>

#include <atomic>

#include <string.h>

#include "common.cuh"
#include "ell_gather.cuh"
#include "gemm.cuh"

namespace cgc {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- tensor-map cache (host).  cuTensorMapEncodeTiled costs microseconds of host time per call and the operand
// views of a training loop repeat every step (the caching allocator hands the same blocks back), so encoded maps
// are kept per thread under a key made of every encode argument.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      return reinterpret_cast<EncodeTiledFn>(sym);
    return nullptr;
  }();
  return fn;
}
bool tensor_map_available() { return encode_fn() != nullptr; }

struct MapKey {
  uint64_t v[10];
};
struct MapSlot {
  MapKey key;
  CUtensorMap map;
  bool used;
};
constexpr int kMapSlots = 64;
static thread_local MapSlot g_maps[kMapSlots];
static thread_local int g_map_next = 0;

static bool cached_encode(const MapKey& key, CUtensorMap* out, cuuint32_t rank, const float* base, const cuuint64_t* dims,
                          const cuuint64_t* strides, const cuuint32_t* box, CUtensorMapSwizzle swz) {
  for (int i = 0; i < kMapSlots; ++i)
    if (g_maps[i].used && memcmp(&g_maps[i].key, &key, sizeof(MapKey)) == 0) {
      *out = g_maps[i].map;
      return true;
    }
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint32_t estr[3] = {1, 1, 1};
  if (fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, const_cast<float*>(base), dims, strides, box, estr,
         CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  MapSlot& s = g_maps[g_map_next];
  g_map_next = (g_map_next + 1) % kMapSlots;
  s.key = key;
  s.map = *out;
  s.used = true;
  return true;
}

bool tensor_map_2d(CUtensorMap* map, const float* base, uint64_t cols, uint64_t rows, int64_t ld, uint32_t box_cols,
                   uint32_t box_rows) {
  MapKey key = {{2, (uint64_t)(uintptr_t)base, cols, rows, (uint64_t)ld, box_cols, box_rows, 0, 0, 0}};
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {box_cols, box_rows};
  return cached_encode(key, map, 2, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
}

bool tensor_map_3d(CUtensorMap* map, const float* base, uint64_t inner, uint64_t rows, uint64_t batch, int64_t ld,
                   int64_t batch_stride, uint32_t box_rows, bool mn_major) {
  MapKey key = {{3, (uint64_t)(uintptr_t)base, inner, rows, batch, (uint64_t)ld, (uint64_t)batch_stride, box_rows,
                 mn_major ? 1u : 0u, 0}};
  cuuint64_t dims[3] = {inner, rows, batch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)(batch > 1 ? batch_stride : (int64_t)rows * ld) * 4};
  cuuint32_t box[3] = {32, box_rows, 1};
  return cached_encode(key, map, 3, base, dims, strides, box,
                       mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B);
}
}  // namespace cgc

using namespace cgc;

extern "C" int cgc_version(void) { return 100; }
extern "C" int64_t cgc_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }
extern "C" const char* cgc_last_error(void) { return g_err; }
extern "C" int64_t cgc_workspace_bytes(void) { return kWorkspaceBytes; }

extern "C" int cgc_gemm(int trans_a, int trans_b, int32_t m, int32_t n, int32_t k, float alpha, const float* a,
                        int64_t lda, int64_t stride_a, const float* b, int64_t ldb, int64_t stride_b, float beta,
                        float* c, int64_t ldc, int64_t stride_c, int32_t batch, const float* bias,
                        const float* row_scale, const int32_t* num_nodes, int32_t nmax, int precision,
                        void* work, int64_t work_bytes, void* stream) {
  CGC_REQUIRE(a && b && c && m > 0 && n > 0 && k > 0 && batch > 0, CGC_ERR_ARG,
              "cgc_gemm: bad args (m=%d n=%d k=%d batch=%d)", m, n, k, batch);
  CGC_REQUIRE(lda >= (trans_a ? m : k) && ldb >= (trans_b ? k : n) && ldc >= n, CGC_ERR_ARG,
              "cgc_gemm: leading dimension too small (lda=%lld ldb=%lld ldc=%lld)", (long long)lda,
              (long long)ldb, (long long)ldc);
  CGC_REQUIRE(num_nodes == nullptr || nmax > 0, CGC_ERR_ARG, "cgc_gemm: mask without nmax");
  CGC_REQUIRE(precision >= -1 && precision <= 3, CGC_ERR_ARG, "cgc_gemm: precision must be -1, 0, 1, 2 or 3");
  GemmParams p;
  p.m = m; p.n = n; p.k = k; p.alpha = alpha; p.beta = beta;
  p.a = a; p.lda = lda; p.stride_a = stride_a;
  p.b = b; p.ldb = ldb; p.stride_b = stride_b;
  p.c = c; p.ldc = ldc; p.stride_c = stride_c;
  p.batch = batch; p.bias = bias; p.row_scale = row_scale; p.num_nodes = num_nodes; p.nmax = nmax;
  p.ksplit = 1; p.kchunk = k; p.partial = nullptr;
  p.a_vec = aligned16(a) && lda % 4 == 0 && stride_a % 4 == 0;
  p.b_vec = aligned16(b) && ldb % 4 == 0 && stride_b % 4 == 0;
  cudaStream_t st = as_stream(stream);
  const bool forced = precision == 1 || precision == 3;
  if (precision != 0 && gemm_tc_supported(p, trans_a, trans_b, forced))
    return gemm_tc(p, trans_a, trans_b, precision >= 2 ? 1 : 3, work, work_bytes, st);
  CGC_REQUIRE(!forced, CGC_ERR_UNSUPPORTED,
              "cgc_gemm: tcgen05 path does not support this shape/alignment (m=%d n=%d k=%d)", m, n, k);
  return gemm_simt(p, trans_a, trans_b, work, work_bytes, st);
}
