// Shared helpers for the knnsvc_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

namespace knnsvc {

void set_error(const char* fmt, ...);

#define KNN_CHECK_ARG(cond, code, ...)                      \
  do {                                                      \
    if (!(cond)) {                                          \
      ::knnsvc::set_error(__VA_ARGS__);                     \
      return (code);                                        \
    }                                                       \
  } while (0)

#define KNN_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      ::knnsvc::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),   \
                          __FILE__, __LINE__);                                      \
      return (int)_e;                                                               \
    }                                                                               \
  } while (0)

void count_launch();
#define KNN_LAUNCH_CHECK()            \
  do {                                \
    ::knnsvc::count_launch();         \
    KNN_CUDA(cudaGetLastError());     \
  } while (0)

// fp16 operand scaling: unit-norm rows are multiplied by 2^10 before the fp16
// cast so that small components stay out of the subnormal range; a dot product
// of two such rows is the cosine similarity times 2^20.
constexpr float kHalfScale = 1024.0f;
constexpr float kDotScale = 1048576.0f;        // 2^20
constexpr float kDotUnscale = 1.0f / 1048576.0f;

// Rigorous half-width of the filter's error window in cosine units:
// |s_fp16gemm - s_exact| <= kFilterEps.  Two fp16 roundings of unit vectors
// contribute <= 2*2^-11 + 2^-22 (Cauchy-Schwarz), the fp32 normalisation and
// <=1024-term fp32 accumulation in the tensor core <= 1.3e-4; see DESIGN.md.
constexpr float kFilterEps = 1.2e-3f;

constexpr int kMaxK = 32;

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace knnsvc
