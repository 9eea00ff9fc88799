// Microbenchmark: issue rate of packed FFMA2 vs scalar FFMA on sm_100a (informs K5/K7 design).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma2_bench tools/micro/ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters) {
  float2 a[8];
  float2 x = make_float2(1.0001f + threadIdx.x * 1e-6f, 0.9999f), y = make_float2(1e-6f, 2e-6f);
  for (int i = 0; i < 8; ++i) a[i] = make_float2(i * 0.1f, i * 0.2f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) {            // 2 scalar FFMA
        a[i].x = fmaf(a[i].x, x.x, y.x);
        a[i].y = fmaf(a[i].y, x.y, y.y);
      } else if (MODE == 1) {     // 1 FFMA2
        a[i] = __ffma2_rn(a[i], x, y);
      } else {                    // 1 FADD2
        a[i] = __fadd2_rn(a[i], y);
      }
    }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
double run(float* out, int blocks, int threads, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(out, iters);
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
  const int iters = 20000;
  for (int threads : {128, 256, 1024}) {
    const int blocks = 148 * (1024 / threads);
    double m0 = run<0>(out, blocks, threads, iters), m1 = run<1>(out, blocks, threads, iters), m2 = run<2>(out, blocks, threads, iters);
    double lanes = (double)blocks * threads * iters * 16;   // fp32 FMA lane-ops
    printf("threads/block %d (warps/SM %d): scalar FFMA %.3f ms (%.1f TFMA/s)  FFMA2 %.3f ms (%.1f TFMA/s)  FADD2 %.3f ms\n", threads,
           1024 / 32, m0, lanes / m0 / 1e9, m1, lanes / m1 / 1e9, m2);
  }
  return 0;
}
