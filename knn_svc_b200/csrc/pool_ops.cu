// Pool-builder tensor ops and the amplitude ratio of the offline
// prematch (rank 2): the pure tensor arithmetic between "WavLM layer features + audio + f0"
// and the matcher's pools.  Reference: get_complete_spk_pool, ddsp_prematch_dataset.py:301-423
// (layer mix :349-350, magnitude STFT :326/:361-363, harmonic amplitudes :391-404) and
// per_spk_extract's amp_ratio :1672-1675.  WavLM itself, file IO and pyworld stay outside.
#include <math.h>

#include "common.cuh"
#include "kernels.cuh"

namespace knnsvc {

// ------------------------------------------------------------------ layer mix
// out[t,d] = sum_l w[l] * feats[l,t,d]   (`(feats*weights[:, None]).sum(dim=0)`, :349-350).
// The reference's weights are float64, so the sum is accumulated in fp64 and
// rounded once.  Both mixes (matching and synthesis weights) come out of ONE pass over the
// layer stack: HBM-bound, L*D*4 bytes read + 2*D*4 written per frame.
struct LayerWeights {
  double a[kMaxLayers];
  double b[kMaxLayers];
};

template <bool TWO>
__global__ void __launch_bounds__(256) layer_mix_kernel(const float* __restrict__ feats, int n_layers, int64_t n_vec,
                                                        LayerWeights w, float* __restrict__ out_a,
                                                        float* __restrict__ out_b) {
  // n_vec = frames*dim/4 float4 elements per layer
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_vec; e += (int64_t)gridDim.x * blockDim.x) {
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0, b0 = 0, b1 = 0, b2 = 0, b3 = 0;
    const float4* src = reinterpret_cast<const float4*>(feats) + e;
#pragma unroll 5
    for (int l = 0; l < n_layers; ++l) {
      const float4 v = __ldcs(src + (int64_t)l * n_vec);   // streamed once
      const double wa = w.a[l];
      a0 += wa * v.x; a1 += wa * v.y; a2 += wa * v.z; a3 += wa * v.w;
      if (TWO) {
        const double wb = w.b[l];
        b0 += wb * v.x; b1 += wb * v.y; b2 += wb * v.z; b3 += wb * v.w;
      }
    }
    reinterpret_cast<float4*>(out_a)[e] = make_float4((float)a0, (float)a1, (float)a2, (float)a3);
    if (TWO) reinterpret_cast<float4*>(out_b)[e] = make_float4((float)b0, (float)b1, (float)b2, (float)b3);
  }
}

int launch_layer_mix(const float* feats, int n_layers, int64_t frames, int dim, const double* w_a_host,
                     const double* w_b_host, float* out_a, float* out_b, cudaStream_t stream) {
  KNN_CHECK_ARG(n_layers >= 1 && n_layers <= kMaxLayers, -3, "layer_mix: %d layers outside [1,%d]", n_layers, kMaxLayers);
  KNN_CHECK_ARG(dim % 4 == 0, -3, "layer_mix: dim %d must be a multiple of 4", dim);
  KNN_CHECK_ARG((reinterpret_cast<uintptr_t>(feats) & 15) == 0 && (reinterpret_cast<uintptr_t>(out_a) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(out_b) & 15) == 0,
                -3, "layer_mix: pointers must be 16-byte aligned");
  if (frames == 0) return 0;
  LayerWeights w;
  for (int l = 0; l < kMaxLayers; ++l) {
    w.a[l] = l < n_layers ? w_a_host[l] : 0.0;
    w.b[l] = (l < n_layers && w_b_host) ? w_b_host[l] : 0.0;
  }
  const int64_t n_vec = frames * dim / 4;
  int64_t grid = ceil_div64(n_vec, 256);
  if (grid > 148 * 16) grid = 148 * 16;
  if (w_b_host)
    layer_mix_kernel<true><<<(unsigned)grid, 256, 0, stream>>>(feats, n_layers, n_vec, w, out_a, out_b);
  else
    layer_mix_kernel<false><<<(unsigned)grid, 256, 0, stream>>>(feats, n_layers, n_vec, w, out_a, nullptr);
  KNN_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------ magnitude STFT
// torchaudio Spectrogram(n_fft, hop, center=True, power=1): periodic Hann window, reflect
// padding, one-sided; the reference drops the Nyquist bin (`.T[:, :-1]`, :361) and keeps the
// first `frames` frames (:363).  Direct DFT: one CTA per ST_FR frames, thread k owns bin k and
// walks the n_fft windowed samples with a twiddle table in shared memory (index k*n mod n_fft
// kept incrementally).  n_fft*n_fft/2 FMAs per frame — a few tens of microseconds per minute
// of audio, so no FFT is needed.
constexpr int ST_FR = 4;
constexpr int ST_MAX_FFT = 1024;

__global__ void __launch_bounds__(256) stft_mag_kernel(const float* __restrict__ x, int64_t n_samples,
                                                       int64_t frames, int n_fft, int hop,
                                                       float* __restrict__ out) {
  __shared__ float s_cos[ST_MAX_FFT], s_sin[ST_MAX_FFT];
  __shared__ float s_x[ST_FR][ST_MAX_FFT];
  const int bins = n_fft / 2;
  const int64_t f0 = (int64_t)blockIdx.x * ST_FR;
  const int pad = n_fft / 2;
  for (int n = threadIdx.x; n < n_fft; n += blockDim.x) {
    double s, c;
    sincospi(2.0 * (double)n / (double)n_fft, &s, &c);
    s_cos[n] = (float)c;
    s_sin[n] = (float)s;
    const float win = (float)(0.5 - 0.5 * c);   // periodic Hann
#pragma unroll
    for (int f = 0; f < ST_FR; ++f) {
      float v = 0.f;
      if (f0 + f < frames) {
        int64_t sidx = (f0 + f) * hop + n - pad;
        if (sidx < 0) sidx = -sidx;                                   // reflect (no edge repeat)
        if (sidx >= n_samples) sidx = 2 * (n_samples - 1) - sidx;
        if (sidx >= 0 && sidx < n_samples) v = __ldg(x + sidx) * win;
      }
      s_x[f][n] = v;
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < bins; k += blockDim.x) {
    float re[ST_FR], im[ST_FR];
#pragma unroll
    for (int f = 0; f < ST_FR; ++f) re[f] = im[f] = 0.f;
    int tw = 0;
    for (int n = 0; n < n_fft; ++n) {
      const float c = s_cos[tw], s = s_sin[tw];
#pragma unroll
      for (int f = 0; f < ST_FR; ++f) {
        const float v = s_x[f][n];
        re[f] = fmaf(v, c, re[f]);
        im[f] = fmaf(v, s, im[f]);
      }
      tw += k;
      if (tw >= n_fft) tw -= n_fft;
    }
#pragma unroll
    for (int f = 0; f < ST_FR; ++f)
      if (f0 + f < frames) out[(f0 + f) * bins + k] = sqrtf(re[f] * re[f] + im[f] * im[f]);
  }
}

int launch_stft_magnitude(const float* audio, int64_t n_samples, int64_t frames, int n_fft, int hop, float* out,
                          cudaStream_t stream) {
  KNN_CHECK_ARG(n_fft >= 2 && n_fft <= ST_MAX_FFT && n_fft % 2 == 0, -3, "stft: n_fft %d outside [2,%d] or odd", n_fft,
                ST_MAX_FFT);
  KNN_CHECK_ARG(hop >= 1 && n_samples > n_fft / 2, -3, "stft: need more than n_fft/2 samples for reflect padding");
  KNN_CHECK_ARG(frames <= 1 + n_samples / hop, -3, "stft: %lld frames requested, the audio has %lld",
                (long long)frames, (long long)(1 + n_samples / hop));
  if (frames == 0) return 0;
  stft_mag_kernel<<<(unsigned)ceil_div64(frames, ST_FR), 256, 0, stream>>>(audio, n_samples, frames, n_fft, hop, out);
  KNN_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------ harmonic amplitudes
// ddsp_prematch_dataset.py:391-404: for harmonic h of frame t the bin
//   round(clamp(f0*h*2*(8S)/sr, max=8S))            (fp32, left to right, half-to-even)
// of the x8 linearly interpolated magnitude row padded with one zero; unvoiced frames keep the
// row maximum in harmonic 1 and zero elsewhere; everything times 0.0108.  The interpolated
// row (F.interpolate scale_factor=8, mode='linear', :395) is never materialised: only the
// <= 49 needed positions are evaluated, with torch's fp32 operation order.
__global__ void __launch_bounds__(256) harmonic_amps_kernel(const float* __restrict__ spec,
                                                            const float* __restrict__ f0, int64_t frames, int n_bins,
                                                            int n_harm, int factor, float sample_rate,
                                                            float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int n_up = n_bins * factor;
  const float inv_factor = 1.0f / (float)factor;
  for (int64_t t = warp; t < frames; t += nwarps) {
    const float f = __ldg(f0 + t);
    const float* row = spec + t * n_bins;
    if (f == 0.f) {
      float mx = -INFINITY;
      for (int c = lane; c < n_bins; c += 32) mx = fmaxf(mx, __ldg(row + c));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      for (int h = lane; h < n_harm; h += 32) out[t * n_harm + h] = h == 0 ? __fmul_rn(0.0108f, mx) : __fmul_rn(0.0108f, 0.f);
      continue;
    }
    for (int h = lane; h < n_harm; h += 32) {
      const float mh = __fmul_rn(f, (float)(h + 1));
      float pos = __fdiv_rn(__fmul_rn(__fmul_rn(mh, 2.0f), (float)n_up), sample_rate);
      pos = fminf(pos, (float)n_up);
      const int bin = (int)rintf(pos);
      float v = 0.f;                                                   // bin == n_up: the F.pad zero
      if (bin < n_up) {
        const float src = fmaxf(__fmul_rn(inv_factor, (float)bin + 0.5f) - 0.5f, 0.f);
        const int i0 = (int)floorf(src);
        const int i1 = i0 + 1 < n_bins ? i0 + 1 : n_bins - 1;
        const float l1 = src - (float)i0, l0 = 1.0f - l1;
        v = __fmaf_rn(l0, __ldg(row + i0), __fmul_rn(l1, __ldg(row + i1)));
      }
      out[t * n_harm + h] = __fmul_rn(0.0108f, v);
    }
  }
}

int launch_harmonic_amplitudes(const float* spec, const float* f0, int64_t frames, int n_bins, int n_harm,
                               int sample_rate, float* out, cudaStream_t stream) {
  KNN_CHECK_ARG(n_bins >= 1 && n_harm >= 1, -3, "harmonic_amplitudes: bad shape");
  if (frames == 0) return 0;
  int64_t grid = ceil_div64(frames, 8);
  if (grid > 148 * 16) grid = 148 * 16;
  harmonic_amps_kernel<<<(unsigned)grid, 256, 0, stream>>>(spec, f0, frames, n_bins, n_harm, 8, (float)sample_rate, out);
  KNN_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------ L1 norms and amp_ratio
// per_spk_extract :1672-1675: amp_ratio[t,k] = |spec_utt[t]|_1 / (|spec_pool[idx[t,k]]|_1 + 1e-5).
// Row norms are computed once per pool (one warp per row, HBM-bound), the ratio is a gather.
__global__ void __launch_bounds__(256) row_l1_kernel(const float* __restrict__ x, int64_t rows, int dim,
                                                     float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = warp; r < rows; r += nwarps) {
    double acc = 0.0;
    for (int c = lane; c < dim; c += 32) acc += (double)fabsf(__ldg(x + r * dim + c));
    acc = warp_sum(acc);
    if (lane == 0) out[r] = (float)acc;
  }
}

int launch_row_l1(const float* x, int64_t rows, int dim, float* out, cudaStream_t stream) {
  if (rows == 0) return 0;
  int64_t grid = ceil_div64(rows, 8);
  if (grid > 148 * 16) grid = 148 * 16;
  row_l1_kernel<<<(unsigned)grid, 256, 0, stream>>>(x, rows, dim, out);
  KNN_LAUNCH_CHECK();
  return 0;
}

__global__ void __launch_bounds__(256) amp_ratio_kernel(const float* __restrict__ l1_query,
                                                        const float* __restrict__ l1_pool,
                                                        const int64_t* __restrict__ idx, int64_t n, int k,
                                                        int64_t n_pool, float* __restrict__ out) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = __ldg(idx + e);
    r = r < 0 ? 0 : (r >= n_pool ? n_pool - 1 : r);
    out[e] = __fdiv_rn(__ldg(l1_query + e / k), __fadd_rn(__ldg(l1_pool + r), 1e-5f));
  }
}

int launch_amp_ratio(const float* l1_query, const float* l1_pool, const int64_t* idx, int64_t n_query, int k,
                     int64_t n_pool, float* out, cudaStream_t stream) {
  if (n_query == 0) return 0;
  const int64_t n = n_query * k;
  int64_t grid = ceil_div64(n, 256);
  if (grid > 148 * 8) grid = 148 * 8;
  amp_ratio_kernel<<<(unsigned)grid, 256, 0, stream>>>(l1_query, l1_pool, idx, n, k, n_pool, out);
  KNN_LAUNCH_CHECK();
  return 0;
}

}  // namespace knnsvc
