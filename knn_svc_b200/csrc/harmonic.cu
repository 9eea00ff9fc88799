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
// Bytes: 200 B in + 1280 B out per frame; the arithmetic (H sines and H 4-tap
// interpolations per 4-byte sample) is what bounds it, see DESIGN.md.
//
// Two kernels:
//   harmonic_group_kernel  (the fast path, H <= 64): one CTA per "tap group" — the hop
//     samples that share the same four bicubic tap frames (second half of frame g-1 and
//     first half of frame g).  The CTA stages the six candidate tap rows in shared
//     memory once; every thread then evaluates ONE accurate sincosf per sample and
//     generates sin(h*phi) for h = 1..H with a plane rotation by 2*phi on packed
//     (odd, even) harmonic pairs — Blackwell's FFMA2 does both lanes per issue slot —
//     instead of H range-reduced sinf calls.  The rotation's error after H/2 steps
//     (<= 2.5e-6) is below the reference's own argument rounding fl(h*phi) (<= 7.6e-6).
//   harmonic_bank_kernel   (general path: single sinusoid, H > 64, or a thread whose
//     fp32 source coordinate falls outside the staged rows): one sinf per harmonic.
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

// bicubic source frame and tap coefficients of output sample n (torch upsample_bicubic2d,
// align_corners=False: fp32 coordinate, clamped lambda)
__device__ __forceinline__ int64_t bicubic_taps(int64_t n, int64_t frames, float scale, float (&c)[4]) {
  const float real = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)n, 0.5f)), 0.5f);
  const float fl = floorf(real);
  int64_t i0 = (int64_t)fl;
  if (i0 > frames - 1) i0 = frames - 1;
  float lam = __fsub_rn(real, (float)i0);
  lam = fminf(fmaxf(lam, 0.0f), 1.0f);
  c[0] = cubic2(__fadd_rn(lam, 1.0f));
  c[1] = cubic1(lam);
  const float x2 = __fsub_rn(1.0f, lam);
  c[2] = cubic1(x2);
  c[3] = cubic2(__fadd_rn(x2, 1.0f));
  return i0;
}

__device__ __forceinline__ int64_t clamp_row(int64_t r, int64_t frames) {
  return r < 0 ? 0 : (r > frames - 1 ? frames - 1 : r);
}

// wrapped fp32 phase of sample j of frame t: fp64 running sum, one cycle removed, THEN cast (:194-195)
__device__ __forceinline__ float wrapped_phase(double base_t, int j, float f0t, int sample_rate) {
  const double two_pi = 6.283185307179586476925286766559;
  const double p = base_t + (double)(j + 1) * ((double)f0t / (double)sample_rate);
  return (float)(two_pi * (p - rint(p)));
}

// one output sample, one range-reduced sinf per harmonic (the reference's operation order)
__device__ __noinline__ float sample_general(const float* __restrict__ am, int64_t frames, int n_harm, float scale,
                                             float nyq, int64_t n, float f0t, float phi) {
  float c[4];
  const int64_t i0 = bicubic_taps(n, frames, scale, c);
  const float* a0 = am + clamp_row(i0 - 1, frames) * n_harm;
  const float* a1 = am + clamp_row(i0, frames) * n_harm;
  const float* a2 = am + clamp_row(i0 + 1, frames) * n_harm;
  const float* a3 = am + clamp_row(i0 + 2, frames) * n_harm;
  float acc = 0.0f;
  for (int h = 1; h <= n_harm; ++h) {
    const float hf = (float)h;
    float a = __fmul_rn(__ldg(a0 + h - 1), c[0]);
    a = __fadd_rn(a, __fmul_rn(__ldg(a1 + h - 1), c[1]));
    a = __fadd_rn(a, __fmul_rn(__ldg(a2 + h - 1), c[2]));
    a = __fadd_rn(a, __fmul_rn(__ldg(a3 + h - 1), c[3]));
    const float mask = (__fmul_rn(f0t, hf) < nyq ? 1.0f : 0.0f) + 1e-7f;
    acc += sinf(__fmul_rn(phi, hf)) * __fmul_rn(a, mask);
  }
  return acc;
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
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < n_samples;
       n += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = n / hop;
    const int j = (int)(n - t * hop);
    const float f0t = __ldg(f + t);
    const float phi = wrapped_phase(bs[t], j, f0t, sample_rate);
    o[n] = HAS_AMP ? sample_general(am, frames, n_harm, scale, nyq, n, f0t, phi) : sinf(phi);
  }
}

// ---- fast path: tap-group CTAs, packed harmonic pairs, rotation recurrence
constexpr int HG_THREADS = 160;
constexpr int HG_ROWS = 6;     // staged amplitude frames g-3 .. g+2 (clamped): covers i0 in {g-2, g-1, g}
constexpr int HG_HMAX = 64;

__device__ __forceinline__ float2 dup2(float v) { return make_float2(v, v); }

// MODE 0: all four harmonics of the chunk are below Nyquist; 1: none is; 2: mixed (per-lane select)
template <int MODE>
__device__ __forceinline__ void harmonic_chunk(const float* __restrict__ rows, int hp, int c, int hcut,
                                               const float2 (&cd)[4], float2 c2d, float2 s2d, float2 ns2d, float2& S,
                                               float2& C, float2& acc_lo, float2& acc_hi) {
  float4 A[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) A[k] = *reinterpret_cast<const float4*>(rows + k * hp + 4 * c);
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    float2 a = __fmul2_rn(half ? make_float2(A[0].z, A[0].w) : make_float2(A[0].x, A[0].y), cd[0]);
#pragma unroll
    for (int k = 1; k < 4; ++k)
      a = __ffma2_rn(half ? make_float2(A[k].z, A[k].w) : make_float2(A[k].x, A[k].y), cd[k], a);
    if (MODE == 0) {
      acc_lo = __ffma2_rn(S, a, acc_lo);
    } else if (MODE == 1) {
      acc_hi = __ffma2_rn(S, a, acc_hi);
    } else {
      const int h0 = 4 * c + 2 * half + 1;  // harmonic number of lane x
      const float2 term = __fmul2_rn(S, a);
      if (h0 <= hcut) acc_lo.x += term.x; else acc_hi.x += term.x;
      if (h0 + 1 <= hcut) acc_lo.y += term.y; else acc_hi.y += term.y;
    }
    // advance both lanes by two harmonics: rotation by 2*phi
    const float2 Sn = __ffma2_rn(S, c2d, __fmul2_rn(C, s2d));
    C = __ffma2_rn(C, c2d, __fmul2_rn(S, ns2d));
    S = Sn;
  }
}

__global__ void __launch_bounds__(HG_THREADS) harmonic_group_kernel(const float* __restrict__ f0,
                                                                    const float* __restrict__ amp, int64_t frames,
                                                                    int n_harm, int sample_rate, int hop,
                                                                    const double* __restrict__ base,
                                                                    float* __restrict__ out) {
  __shared__ __align__(16) float s_rows[HG_ROWS * HG_HMAX];
  const int b = blockIdx.y;
  const int64_t g = blockIdx.x;                        // tap group: i0 == g - 1 for (almost) all of its samples
  const int64_t n_samples = frames * hop;
  const float* f = f0 + (int64_t)b * frames;
  const double* bs = base + (int64_t)b * frames;
  const float* am = amp + (int64_t)b * frames * n_harm;
  float* o = out + (int64_t)b * n_samples;
  const float scale = (float)frames / (float)n_samples;
  const float nyq = (float)sample_rate / 2.0f;
  const int hp = (n_harm + 3) & ~3;                    // padded harmonics carry zero amplitude

  for (int e = threadIdx.x; e < HG_ROWS * hp; e += HG_THREADS) {
    const int r = e / hp, h = e - r * hp;
    s_rows[e] = h < n_harm ? __ldg(am + clamp_row(g - 3 + r, frames) * n_harm + h) : 0.0f;
  }
  __syncthreads();

  const int64_t n_begin = g * hop - hop / 2;
  for (int m = threadIdx.x; m < hop; m += HG_THREADS) {
    const int64_t n = n_begin + m;
    if (n < 0 || n >= n_samples) continue;
    // the group is the tail of frame g-1 followed by the head of frame g: no division needed
    const int head = hop - hop / 2;                    // samples of frame g-1 in this group
    const int64_t t = m < hop / 2 ? g - 1 : g;
    const int j = m < hop / 2 ? m + head : m - hop / 2;
    const float f0t = __ldg(f + t);
    const float phi = wrapped_phase(bs[t], j, f0t, sample_rate);
    float c[4];
    const int64_t i0 = bicubic_taps(n, frames, scale, c);
    const int64_t rbase = i0 - g + 2;                  // staged row of tap frame i0 - 1
    if (rbase < 0 || rbase > HG_ROWS - 4) {            // fp32 coordinate rounding on very long inputs
      o[n] = sample_general(am, frames, n_harm, scale, nyq, n, f0t, phi);
      continue;
    }
    // number of leading harmonics with fl(f0*h) < sr/2 (the product is monotone in h) — :146-156
    int hcut = n_harm;
    if (!(__fmul_rn(f0t, (float)n_harm) < nyq)) {
      int e = (int)floorf(nyq / f0t);
      e = e < 0 ? 0 : (e > n_harm ? n_harm : e);
      while (e > 0 && !(__fmul_rn(f0t, (float)e) < nyq)) --e;
      while (e < n_harm && __fmul_rn(f0t, (float)(e + 1)) < nyq) ++e;
      hcut = e;
    }
    float s1, c1;
    sincosf(phi, &s1, &c1);
    const float s2 = 2.0f * s1 * c1, c2 = fmaf(-2.0f * s1, s1, 1.0f);
    float2 S = make_float2(s1, s2), C = make_float2(c1, c2);
    const float2 c2d = dup2(c2), s2d = dup2(s2), ns2d = dup2(-s2);
    const float2 cd[4] = {dup2(c[0]), dup2(c[1]), dup2(c[2]), dup2(c[3])};
    float2 acc_lo = make_float2(0.f, 0.f), acc_hi = make_float2(0.f, 0.f);
    const float* rows = s_rows + rbase * hp;
    const int n_chunks = hp / 4, lo_chunks = hcut / 4;
    int ch = 0;
    for (; ch < lo_chunks; ++ch) harmonic_chunk<0>(rows, hp, ch, hcut, cd, c2d, s2d, ns2d, S, C, acc_lo, acc_hi);
    if (ch < n_chunks && (hcut & 3)) {
      harmonic_chunk<2>(rows, hp, ch, hcut, cd, c2d, s2d, ns2d, S, C, acc_lo, acc_hi);
      ++ch;
    }
    for (; ch < n_chunks; ++ch) harmonic_chunk<1>(rows, hp, ch, hcut, cd, c2d, s2d, ns2d, S, C, acc_lo, acc_hi);
    // amplitudes carry the factor (h*f0 < sr/2) + 1e-7 (:153)
    o[n] = (acc_lo.x + acc_lo.y) * (1.0f + 1e-7f) + (acc_hi.x + acc_hi.y) * 1e-7f;
  }
}

int launch_harmonic_bank(const float* f0, const float* amp, int batch, int64_t frames, int n_harm, int sample_rate,
                         int hop, float* out, double* phase_ws, cudaStream_t stream) {
  if (batch == 0 || frames == 0) return 0;
  KNN_CHECK_ARG(hop >= 1 && sample_rate >= 1, -3, "harmonic_bank: bad hop/sample_rate");
  KNN_CHECK_ARG(batch <= 65535, -3, "harmonic_bank: batch %d exceeds 65535", batch);
  phase_scan_kernel<<<batch, HS_THREADS, 0, stream>>>(f0, frames, sample_rate, hop, phase_ws);
  KNN_LAUNCH_CHECK();
  const int64_t n_samples = frames * hop;
  if (amp && n_harm >= 1 && n_harm <= HG_HMAX && frames + 1 < ((int64_t)1 << 31)) {
    dim3 grid((unsigned)(frames + 1), batch);
    harmonic_group_kernel<<<grid, HG_THREADS, 0, stream>>>(f0, amp, frames, n_harm, sample_rate, hop, phase_ws, out);
    KNN_LAUNCH_CHECK();
    return 0;
  }
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
