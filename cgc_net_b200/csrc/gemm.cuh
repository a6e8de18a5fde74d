// Shared declarations of the two GEMM back ends (fp32 FMA and tcgen05 3xTF32).
#pragma once
#include "common.cuh"

namespace cgc {

struct GemmParams {
  int m, n, k;
  float alpha, beta;
  const float* a; int64_t lda, stride_a;
  const float* b; int64_t ldb, stride_b;
  float* c; int64_t ldc, stride_c;
  int batch;
  const float* bias;          // [n] or null
  const float* row_scale;     // [batch*m] or null
  const int32_t* num_nodes;   // row mask over the batch*m padded rows, or null
  int nmax;
  int ksplit, kchunk;         // split-K (fp32 FMA path)
  float* partial;
  bool a_vec, b_vec;          // 16-byte vector loads allowed on A / B
};

// final write of one output element; v already includes alpha
__device__ __forceinline__ void gemm_store(const GemmParams& p, int bz, int row, int col, float v) {
  const int64_t grow = (int64_t)bz * p.m + row;
  if (p.bias) v += __ldg(p.bias + col);
  if (p.row_scale) v *= __ldg(p.row_scale + grow);
  if (p.num_nodes && !row_valid(p.num_nodes, p.nmax, grow)) v = 0.f;
  float* dst = p.c + (int64_t)bz * p.stride_c + (int64_t)row * p.ldc + col;
  if (p.beta != 0.f) v += p.beta * (*dst);
  *dst = v;
}

int gemm_simt(GemmParams p, int ta, int tb, void* work, int64_t work_bytes, cudaStream_t st);

// tcgen05 3xTF32 path.  Returns 1 ("not taken") when the shape / alignment is unsupported.
bool gemm_tc_supported(const GemmParams& p, int ta, int tb, bool forced);
// passes: 3 = 3xTF32 (fp32-grade products), 1 = one TF32 pass on round-to-nearest operands
int gemm_tc(GemmParams p, int ta, int tb, int passes, void* work, int64_t work_bytes, cudaStream_t st);

}  // namespace cgc
