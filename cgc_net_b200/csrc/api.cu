// Library-level entry points: version, error string, workspace size, and the cgc_gemm dispatcher.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"
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
  CGC_REQUIRE(precision >= -1 && precision <= 2, CGC_ERR_ARG, "cgc_gemm: precision must be -1, 0, 1 or 2");
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
  if (precision != 0 && gemm_tc_supported(p, trans_a, trans_b, precision == 1))
    return gemm_tc(p, trans_a, trans_b, precision == 2 ? 1 : 3, work, work_bytes, st);
  CGC_REQUIRE(precision != 1, CGC_ERR_UNSUPPORTED,
              "cgc_gemm: tcgen05 path does not support this shape/alignment (m=%d n=%d k=%d)", m, n, k);
  return gemm_simt(p, trans_a, trans_b, work, work_bytes, st);
}
