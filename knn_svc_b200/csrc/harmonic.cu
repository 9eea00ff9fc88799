// K7 / K7': additive harmonic bank that conditions the vocoder.
//   get_bulk_dsp_choral(f0, amp, 16000, 320)  — ddsp_prematch_dataset.py:165-208
//   single f0 sinusoid                         — hifigan/ddsp_models_f0.py:344-352
//
// The reference materialises three [B, 320T, H] tensors; here one fused kernel
// produces each output sample from 4 amplitude frames and one phase:
//   phase   : fp64 running sum of f0_up/sr (cumsum at :194).  f0 is upsampled by
//             nearest (x hop), so inside frame t the sum is base[t] + (j+1)*f0[t]/sr;
//             base[] is an fp64 exclusive scan over frames (one CTA per batch row).
//   wrap    : 2*pi*(p - rint(p)) cast to fp32, THEN multiplied by h in fp32 (:195-196).
//   amp_up  : bicubic x hop, A=-0.75, half-pixel centres, border-clamped taps,
//             evaluated in fp32 with torch's operation order (upsample_bicubic2d).
//   out     : sum_h sinf(h*phi) * amp_up_h * ((h*f0_up < sr/2) + 1e-7)   (:146-156, :206)
// Bytes: 200 B in + 1280 B out per frame; the kernel is fp32-ALU/SFU bound
// (49 accurate sinf per sample), see DESIGN.md.
#include "common.cuh"
#include "kernels.cuh"

namespace knnsvc {

// ---- fp64 exclusive scan of hop*f0[t]/sr over frames; one CTA per batch row
constexpr int HS_THREADS = 1024;

__global__ void __launch_bounds__(HS_THREADS) phase_scan_kernel(const float* __restrict__ f0, int64_t frames,
                                                                int sample_rate, int hop,
                                                                double* __restrict__ base) {
  const int b = blockIdx.x;
  const float* f = f0 + (int64_t)b * frames;
  double* o = base + (int64_t)b * frames;
  __shared__ double s_part[HS_THREADS];
  const int tid = threadIdx.x;
  const int64_t per = ceil_div64(frames, HS_THREADS);
  const int64_t a = (int64_t)tid * per, e = min(frames, a + per);
  double sum = 0.0;
  for (int64_t t = a; t < e; ++t) sum += (double)hop * ((double)f[t] / (double)sample_rate);
  s_part[tid] = sum;
  __syncthreads();
  // inclusive Hillis-Steele scan over the 1024 partials
  for (int off = 1; off < HS_THREADS; off <<= 1) {
    double v = (tid >= off) ? s_part[tid - off] : 0.0;
    __syncthreads();
    s_part[tid] += v;
    __syncthreads();
  }
  double run = (tid == 0) ? 0.0 : s_part[tid - 1];
  for (int64_t t = a; t < e; ++t) {
    o[t] = run;
    run += (double)hop * ((double)f[t] / (double)sample_rate);
  }
}

// torch's cubic convolution coefficients (A = -0.75), fp32, unfused like the CPU build
__device__ __forceinline__ float cubic1(float x) {  // |x| <= 1
  const float A = -0.75f;
  return __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(A + 2.0f, x), A + 3.0f), x), x), 1.0f);
}
__device__ __forceinline__ float cubic2(float x) {  // 1 < |x| < 2
  const float A = -0.75f;
  return __fsub_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(__fmul_rn(A, x), 5.0f * A), x), 8.0f * A), x), 4.0f * A);
}

template <bool HAS_AMP>
__global__ void __launch_bounds__(256) harmonic_bank_kernel(const float* __restrict__ f0,
                                                            const float* __restrict__ amp, int64_t frames,
                                                            int n_harm, int sample_rate, int hop,
                                                            const double* __restrict__ base,
                                                            float* __restrict__ out) {
  const int b = blockIdx.y;
  const int64_t n_samples = frames * hop;
  const float* f = f0 + (int64_t)b * frames;
  const double* bs = base + (int64_t)b * frames;
  const float* am = HAS_AMP ? amp + (int64_t)b * frames * n_harm : nullptr;
  float* o = out + (int64_t)b * n_samples;
  const float scale = (float)frames / (float)n_samples;  // area_pixel_compute_scale, align_corners=False
  const float nyq = (float)sample_rate / 2.0f;
  const double two_pi = 6.283185307179586476925286766559;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < n_samples;
       n += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = n / hop;
    const int j = (int)(n - t * hop);
    const float f0t = __ldg(f + t);
    const double p = bs[t] + (double)(j + 1) * ((double)f0t / (double)sample_rate);
    const float phi = (float)(two_pi * (p - rint(p)));
    if (!HAS_AMP) {
      o[n] = sinf(phi);
      continue;
    }
    // bicubic source coordinate and taps
    const float real = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)n, 0.5f)), 0.5f);
    const float fl = floorf(real);
    int64_t i0 = (int64_t)fl;
    if (i0 > frames - 1) i0 = frames - 1;
    float lam = __fsub_rn(real, (float)i0);
    lam = fminf(fmaxf(lam, 0.0f), 1.0f);
    const float c0 = cubic2(__fadd_rn(lam, 1.0f));
    const float c1 = cubic1(lam);
    const float x2 = __fsub_rn(1.0f, lam);
    const float c2 = cubic1(x2);
    const float c3 = cubic2(__fadd_rn(x2, 1.0f));
    int64_t r0 = i0 - 1, r1 = i0, r2 = i0 + 1, r3 = i0 + 2;
    r0 = r0 < 0 ? 0 : (r0 > frames - 1 ? frames - 1 : r0);
    r1 = r1 < 0 ? 0 : (r1 > frames - 1 ? frames - 1 : r1);
    r2 = r2 < 0 ? 0 : (r2 > frames - 1 ? frames - 1 : r2);
    r3 = r3 < 0 ? 0 : (r3 > frames - 1 ? frames - 1 : r3);
    const float* a0 = am + r0 * n_harm;
    const float* a1 = am + r1 * n_harm;
    const float* a2 = am + r2 * n_harm;
    const float* a3 = am + r3 * n_harm;
    float acc = 0.0f;
    for (int h = 1; h <= n_harm; ++h) {
      const float hf = (float)h;
      float a = __fmul_rn(__ldg(a0 + h - 1), c0);
      a = __fadd_rn(a, __fmul_rn(__ldg(a1 + h - 1), c1));
      a = __fadd_rn(a, __fmul_rn(__ldg(a2 + h - 1), c2));
      a = __fadd_rn(a, __fmul_rn(__ldg(a3 + h - 1), c3));
      const float mask = (__fmul_rn(f0t, hf) < nyq ? 1.0f : 0.0f) + 1e-7f;
      acc += sinf(__fmul_rn(phi, hf)) * __fmul_rn(a, mask);
    }
    o[n] = acc;
  }
}

int launch_harmonic_bank(const float* f0, const float* amp, int batch, int64_t frames, int n_harm, int sample_rate,
                         int hop, float* out, double* phase_ws, cudaStream_t stream) {
  if (batch == 0 || frames == 0) return 0;
  KNN_CHECK_ARG(hop >= 1 && sample_rate >= 1, -3, "harmonic_bank: bad hop/sample_rate");
  phase_scan_kernel<<<batch, HS_THREADS, 0, stream>>>(f0, frames, sample_rate, hop, phase_ws);
  KNN_LAUNCH_CHECK();
  const int64_t n_samples = frames * hop;
  int64_t gx = ceil_div64(n_samples, 256);
  if (gx > 148 * 32) gx = 148 * 32;
  dim3 grid((unsigned)gx, batch);
  if (amp)
    harmonic_bank_kernel<true><<<grid, 256, 0, stream>>>(f0, amp, frames, n_harm, sample_rate, hop, phase_ws, out);
  else
    harmonic_bank_kernel<false><<<grid, 256, 0, stream>>>(f0, amp, frames, n_harm, sample_rate, hop, phase_ws, out);
  KNN_LAUNCH_CHECK();
  return 0;
}

}  // namespace knnsvc
