// Shared helpers for the knnsvc_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include <atomic>

namespace knnsvc {

// Function attributes (cudaFuncSetAttribute), SM counts and occupancy answers belong to a DEVICE,
// not to the process: every cache of one is an array indexed by the current device.  slot() is
// nullptr when the device index is unknown or out of range — the caller then does the work every time.
constexpr int kMaxDevices = 64;
struct PerDevice {
  std::atomic<int> v[kMaxDevices];
  std::atomic<int>* slot() {
    int d = -1;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= kMaxDevices) return nullptr;
    return &v[d];
  }
};
// Opt a kernel in to `bytes` of dynamic shared memory on the current device (once per device and size).
#define KNN_SMEM_ATTR(cache, func, bytes)                                                                   \
  do {                                                                                                      \
    std::atomic<int>* _s = (cache).slot();                                                                  \
    if (!_s || _s->load(std::memory_order_acquire) < (int)(bytes)) {                                        \
      KNN_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));      \
      if (_s) _s->store((int)(bytes), std::memory_order_release);                                           \
    }                                                                                                       \
  } while (0)

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

// Rigorous half-width eps of the filter's error window in cosine units:
// |s_fp16gemm - s_exact| <= eps = rho_q + rho_p + rho_q*rho_p + acc_eps(dim), where
// rho = |u - x/|x||_2 is the distance between a row's fp16 operand u (unscaled) and its exact
// unit vector (Cauchy-Schwarz on  u.v - q.p = (u-q).v + q.(v-p)), and acc_eps bounds what the
// tensor core's fp32 accumulation adds: dim/16 MMAs of K = 16 each fold 16 exact products into the
// fp32 accumulator (|partial sums| <= ~1; truncation allowed) — 1.4e-4 per 1024 dimensions, i.e.
// 2^-23 for each of ~1100 additions.  tests/test_gpu_parity.py reads the accumulator values back
// and OBSERVES both parts (test_filter_accumulator_error_is_inside_the_window).
// prepare_rows MEASURES rho per row and publishes the maximum over the row set, so eps is as
// tight as the data allows (~6e-4 on Gaussian-like rows); without a measured value the worst
// case of an fp16 rounding, 2^-11 (+ fp32 normalisation), is assumed: 2*5.3e-4 + 1.4e-4 = 1.2e-3.
constexpr float kAccEps = 1.4e-4f;        // per 1024 (padded) feature dimensions
constexpr float kWorstRowEps = 5.3e-4f;
__host__ __device__ inline float filter_acc_eps(int dim_pad) {
  return dim_pad <= 1024 ? kAccEps : kAccEps * ((float)dim_pad * (1.0f / 1024.0f));
}
__host__ __device__ inline float filter_eps_from(float rho_q, float rho_p, int dim_pad) {
  return rho_q + rho_p + rho_q * rho_p + filter_acc_eps(dim_pad);
}
// rho as published by prepare_rows (device scalars; nullptr = worst case)
__device__ __forceinline__ float filter_eps(const float* __restrict__ q_err, const float* __restrict__ p_err,
                                            int dim_pad) {
  const float rq = q_err ? __ldg(q_err) : kWorstRowEps;
  const float rp = p_err ? __ldg(p_err) : kWorstRowEps;
  return filter_eps_from(rq, rp, dim_pad);
}

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
