// Shared helpers of libcgc_sm100 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/cgc_sm100.h"

namespace cgc {

// thread-local error message returned by cgc_last_error()
void set_error(const char* fmt, ...);
// process-wide count of kernels this library has launched (cgc_launch_count())
void count_launch();

#define CGC_REQUIRE(cond, code, ...)      \
  do {                                    \
    if (!(cond)) {                        \
      ::cgc::set_error(__VA_ARGS__);      \
      return (code);                      \
    }                                     \
  } while (0)

// Check the launch of the kernel just enqueued (no synchronisation).
#define CGC_CHECK_LAUNCH(name)                                                  \
  do {                                                                          \
    cudaError_t e_ = cudaGetLastError();                                        \
    if (e_ != cudaSuccess) {                                                    \
      ::cgc::set_error("%s: %s", name, cudaGetErrorString(e_));                 \
      return (int)e_;                                                           \
    }                                                                           \
    ::cgc::count_launch();                                                      \
  } while (0)

#define CGC_CUDA(call)                                                          \
  do {                                                                          \
    cudaError_t e_ = (call);                                                    \
    if (e_ != cudaSuccess) {                                                    \
      ::cgc::set_error("%s: %s", #call, cudaGetErrorString(e_));                \
      return (int)e_;                                                           \
    }                                                                           \
  } while (0)

constexpr int kNumSMs = 148;            // B200: 2 dies x 74 SMs
constexpr float kL2NormEps = 1e-12f;    // F.normalize eps
constexpr float kRenormEps = 1e-15f;    // model/network.py:8

// Row mask in the padded node layout: row r is a real node iff r % nmax < num_nodes[r / nmax].
// Padded row counts fit 32 bits everywhere in the library (the CSR indices are int32), so the division is a
// 32-bit one (a 64-bit division costs ~3x the instructions) and the function stays branch-free: it is inlined
// into load loops whose batching a branch would break.
__device__ __forceinline__ bool row_valid(const int32_t* __restrict__ num_nodes, int nmax, int64_t r) {
  if (num_nodes == nullptr) return true;
  const uint32_t ru = (uint32_t)r;
  const uint32_t g = ru / (uint32_t)nmax;
  return (int)(ru - g * (uint32_t)nmax) < __ldg(num_nodes + g);
}

// Activation that follows a conv (GNN_Module._activation, network.py:82-90): 0 relu, 1 elu (alpha 1), 2 leaky relu
// (slope 0.01).  The gradient is written in terms of the OUTPUT y, as the reference's in-place modules do.
__device__ __forceinline__ float act_fwd(int act, float v) {
  if (act == CGC_ACT_RELU) return fmaxf(v, 0.f);
  if (act == CGC_ACT_ELU) return v > 0.f ? v : expm1f(v);
  return v > 0.f ? v : 0.01f * v;
}
__device__ __forceinline__ float act_grad(int act, float y) {
  if (act == CGC_ACT_RELU) return y > 0.f ? 1.f : 0.f;
  if (act == CGC_ACT_ELU) return y > 0.f ? 1.f : y + 1.f;
  return y > 0.f ? 1.f : 0.01f;
}
// Branch-free form for the hot backward loops: act'(y) = y > 0 ? 1 : a0 + a1 * y with (a0, a1) = (0, 0) relu,
// (1, 1) elu, (0.01, 0) leaky relu -- one select + one FMA per element, no control flow between the loads.
struct ActGrad {
  float a0, a1;
  __device__ __forceinline__ explicit ActGrad(int act)
      : a0(act == CGC_ACT_RELU ? 0.f : (act == CGC_ACT_ELU ? 1.f : 0.01f)), a1(act == CGC_ACT_ELU ? 1.f : 0.f) {}
  __device__ __forceinline__ float operator()(float y) const { return y > 0.f ? 1.f : fmaf(a1, y, a0); }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Tuning / profiling switches read from the environment exist only in lab builds (make LAB=1 -> -DCGC_LAB): the
// release library never lets an environment variable change its numerics or skip work.
#ifdef CGC_LAB
inline int lab_env(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
#else
inline int lab_env(const char*, int dflt) { return dflt; }
#endif
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device property of a kernel: raise it only when a launch
// needs more than what has been set on the CURRENT device so far (`state` = one static array per kernel
// instantiation), instead of paying the driver call on every launch.
constexpr int kMaxDevices = 64;
template <typename Kern>
inline cudaError_t ensure_dynamic_smem(Kern kern, size_t smem, size_t* state) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < kMaxDevices && smem <= state[dev]) return cudaSuccess;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess && dev >= 0 && dev < kMaxDevices) state[dev] = smem;
  return e;
}

// Scratch layout: every entry point that needs scratch uses the front of the shared buffer.
constexpr int64_t kWorkspaceBytes = 256ll << 20;   // 256 MiB (split-K partials dominate)
// The LAST kCounterBytes of the workspace hold self-resetting completion counters ("the last CTA finalises"): zero
// when the buffer is first handed to the library, zero again after every call that used them.
constexpr int64_t kCounterBytes = 256;

}  // namespace cgc
