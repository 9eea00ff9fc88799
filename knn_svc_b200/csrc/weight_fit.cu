// K6: Adam(amsgrad) fit of per-frame softmax mixing weights —
// compute_wavlm_weight (ddsp_prematch_dataset.py:574-680, loss scale 0.1) and
// compute_extended_weight (:807-924, loss scale 1000; its scaling_factors are inert).
//
// The reference re-gathers 3x[T,4,D] rows and runs autograd every iteration.
// The loss is quadratic in the softmax weights w:
//   L = c/((T-1)D) * sum_t u_t' G_t u_t,   u_t = (w[t+1,:], -w[t,:])  in R^8,
//   G_t = Gram{S_-1[t+1,k], S_0[t,k]} + Gram{S_0[t+1,k], S_+1[t,k]},  S_i = synth[clamp(idx+i)]
// so the 8x8 Gram blocks are built ONCE (fp64, HBM-bound gather: 3*K*D*4 bytes
// per frame) and the whole optimisation loop then runs inside one persistent
// CTA on [T,4] state with the analytic gradient.  Logits and Adam moments are
// fp32 and the update is torch.optim.Adam(amsgrad=True)'s single-tensor form;
// the loss and dL/dw are fp64 and cast to fp32 at w, as on the reference's real
// (float64) path.  Stop rules: ddsp_prematch_dataset.py:644-663.
#include <cooperative_groups.h>
#include <limits.h>

#include "common.cuh"
#include "kernels.cuh"

namespace knnsvc {

constexpr int WF_K = 4;
constexpr int kWcMaxSmem = 220 * 1024;   // dynamic shared memory the cluster kernel may ask for
constexpr int WF_V = 2 * WF_K;         // vectors per Gram block
// The 8x8 Gram block [[A, B], [B', C]] is symmetric: 10 + 16 + 10 = 36 unique entries, stored
// ENTRY-MAJOR (gram[e * n_pairs + t]) so that consecutive threads (frames) read consecutive
// addresses in the optimisation loop.
constexpr int WF_E = 36;
__host__ __device__ constexpr int wf_tri(int i, int j) { return i * 4 - i * (i - 1) / 2 + (j - i); }  // i <= j < 4
__host__ __device__ constexpr int wf_a(int i, int j) { return i <= j ? wf_tri(i, j) : wf_tri(j, i); }
__host__ __device__ constexpr int wf_b(int i, int j) { return 10 + i * 4 + j; }                        // A-row i, C-col j
__host__ __device__ constexpr int wf_c(int i, int j) { return 26 + (i <= j ? wf_tri(i, j) : wf_tri(j, i)); }
constexpr int WF_THREADS = 512;
constexpr int WF_SMEM_FRAMES = 10240;  // frames whose weights fit in shared memory
constexpr int WF_STATE_FRAMES = 1900;  // frames whose weights AND optimiser state (7 x 16 B per frame) fit

// ---- Gram build: one WARP per frame pair t (frames t, t+1).  For each of the two blocks the lane
// loads its columns of all 8 rows ONCE (8 loads) and forms the 36 unique products of the
// symmetric 8x8 Gram from registers: 36 DFMA per 8 loads, so the kernel runs at the fp64 pipe's
// rate instead of the L1 load-issue rate (the first version re-read the 8 rows in every warp:
// 9 loads per 8 DFMA, 5x slower at D = 1024).  Both blocks accumulate into the same 36 sums
// (G_t is their sum).  Products of two fp32 values are exact in fp64.
constexpr int WG_WARPS = 8;

__device__ __forceinline__ void wg_accumulate(const double (&v)[WF_V], double (&acc)[WF_E]) {
#pragma unroll
  for (int i = 0; i < WF_K; ++i)           // A: rows 0..3 x rows i..3      (entries 0..9)
#pragma unroll
    for (int j = i; j < WF_K; ++j) acc[wf_tri(i, j)] = fma(v[i], v[j], acc[wf_tri(i, j)]);
#pragma unroll
  for (int i = 0; i < WF_K; ++i)           // B: rows 0..3 x rows 4..7      (entries 10..25)
#pragma unroll
    for (int j = 0; j < WF_K; ++j) acc[wf_b(i, j)] = fma(v[i], v[WF_K + j], acc[wf_b(i, j)]);
#pragma unroll
  for (int i = 0; i < WF_K; ++i)           // C: rows 4..7 x rows 4+i..7    (entries 26..35)
#pragma unroll
    for (int j = i; j < WF_K; ++j) acc[26 + wf_tri(i, j)] = fma(v[WF_K + i], v[WF_K + j], acc[26 + wf_tri(i, j)]);
}

__global__ void __launch_bounds__(WG_WARPS * 32) weight_gram_kernel(const int64_t* __restrict__ idx,
                                                                    const __grid_constant__ RowTable synth,
                                                                    int dim, int64_t n_query, double* __restrict__ gram) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t n_pairs = n_query - 1;
  const int64_t t = (int64_t)blockIdx.x * WG_WARPS + warp;
  if (t >= n_pairs) return;
  // rows of the two blocks: block 0 = (S_-1[t+1,k], S_0[t,k]), block 1 = (S_0[t+1,k], S_+1[t,k])
  const float* rows[2][WF_V];
#pragma unroll
  for (int j = 0; j < WF_V; ++j) {
    const int64_t base = (j < WF_K) ? __ldg(idx + (t + 1) * WF_K + j) : __ldg(idx + t * WF_K + (j - WF_K));
    int64_t r0 = base + ((j < WF_K) ? -1 : 0);
    int64_t r1 = base + ((j < WF_K) ? 0 : 1);
    rows[0][j] = table_row(synth, r0, dim);      // clamped to the pool's ends like the reference's gather
    rows[1][j] = table_row(synth, r1, dim);
  }
  double acc[WF_E];
#pragma unroll
  for (int e = 0; e < WF_E; ++e) acc[e] = 0.0;
  const bool vec = (dim & 3) == 0 && table_aligned16(synth);
#pragma unroll
  for (int b = 0; b < 2; ++b) {
    if (vec) {
#pragma unroll 1
      for (int c = lane; c < dim / 4; c += 32) {
        float4 x[WF_V];
#pragma unroll
        for (int j = 0; j < WF_V; ++j) x[j] = __ldg(reinterpret_cast<const float4*>(rows[b][j]) + c);
        double v[WF_V];
#pragma unroll
        for (int j = 0; j < WF_V; ++j) v[j] = (double)x[j].x;
        wg_accumulate(v, acc);
#pragma unroll
        for (int j = 0; j < WF_V; ++j) v[j] = (double)x[j].y;
        wg_accumulate(v, acc);
#pragma unroll
        for (int j = 0; j < WF_V; ++j) v[j] = (double)x[j].z;
        wg_accumulate(v, acc);
#pragma unroll
        for (int j = 0; j < WF_V; ++j) v[j] = (double)x[j].w;
        wg_accumulate(v, acc);
      }
    } else {
#pragma unroll 1
      for (int c = lane; c < dim; c += 32) {
        double v[WF_V];
#pragma unroll
        for (int j = 0; j < WF_V; ++j) v[j] = (double)__ldg(rows[b][j] + c);
        wg_accumulate(v, acc);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < WF_E; ++e) {
    const double v = warp_sum(acc[e]);
    if (lane == (e & 31)) gram[(int64_t)e * n_pairs + t] = v;   // 36 entries over 32 lanes
  }
}

struct WfState {
  float* theta;
  float* m;
  float* v;
  float* vmax;
  float* best;
  float* grad;   // dL/dw scratch
  float* wglob;  // [T,4] softmax weights when they do not fit in shared memory
  double* chunk; // [T/32 + n_utt] loss sums of 32-frame chunks when they do not fit in shared memory
};

__device__ __forceinline__ void softmax4(const float* th, float* w) {
  const float mx = fmaxf(fmaxf(th[0], th[1]), fmaxf(th[2], th[3]));
  float e[4], s = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    e[k] = expf(th[k] - mx);
    s += e[k];
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) w[k] = e[k] / s;
}

// ---- pieces shared by the one-CTA and the cluster kernel, so that both run the same arithmetic
// in the same order (an utterance's result must not depend on which of the two fitted it).

// Frame t's share e_t of the quadratic form and dL/d(a.w) of its four weights.  wt / wp / wn are the
// (amp-scaled) weights of frames t, t-1, t+1 in fp64; gp(e) / gc(e) return entry e of the Gram
// blocks of pairs (t-1, t) and (t, t+1).
template <class GramPrev, class GramCur>
__device__ __forceinline__ double wf_frame(bool has_prev, bool has_next, const double (&wt)[4], const double (&wp)[4],
                                           const double (&wn)[4], GramPrev gp, GramCur gc, double (&g)[4]) {
  double e_t = 0.0;
#pragma unroll
  for (int k = 0; k < 4; ++k) g[k] = 0.0;
  if (has_prev) {  // pair t-1: this frame is the "t+1" member -> rows 0..3 of G u = A w[t] - B w[t-1]
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      double y = 0.0;
#pragma unroll
      for (int c = 0; c < 4; ++c) y += gp(wf_a(r, c)) * wt[c] - gp(wf_b(r, c)) * wp[c];
      e_t += wt[r] * y;
      g[r] += y;
    }
  }
  if (has_next) {  // pair t: this frame is the "t" member -> rows 4..7 of G u = B' w[t+1] - C w[t]
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      double y = 0.0;
#pragma unroll
      for (int c = 0; c < 4; ++c) y += gc(wf_b(c, r)) * wn[c] - gc(wf_c(r, c)) * wt[c];
      e_t -= wt[r] * y;
      g[r] -= y;
    }
  }
  return e_t;
}

// torch.optim.Adam(amsgrad=True), single-tensor form, on the four logits of one frame
__device__ __forceinline__ void wf_adam4(float* theta, float* m_, float* v_, float* vmax_, const float (&w)[4],
                                         const float (&gw)[4], float step_size, float bc2_sqrt) {
  const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
  float dotgw = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) dotgw += gw[k] * w[k];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float gth = w[k] * (gw[k] - dotgw);  // softmax backward
    float m = m_[k], v = v_[k], vm = vmax_[k];
    m = m + (gth - m) * (1.0f - b1);
    v = v * b2 + (1.0f - b2) * gth * gth;
    vm = fmaxf(vm, v);
    const float denom = sqrtf(vm) / bc2_sqrt + eps;
    theta[k] = theta[k] - step_size * (m / denom);
    m_[k] = m;
    v_[k] = v;
    vmax_[k] = vm;
  }
}

// The reference's stop rules (ddsp_prematch_dataset.py:644-663) and Adam's bias corrections, run by
// one thread per CTA.  update() returns bit0: stop, bit1: snapshot the logits.
struct WfControl {
  double min_loss = 20000.0, converge_min_loss = 20000.0, first_loss = 0.0, last_loss = 0.0;
  int since_improve = 0, stop_iter;
  __device__ explicit WfControl(int max_iters) : stop_iter(max_iters) {}
  __device__ int update(double loss, int it) {
    if (it == 0) first_loss = loss;
    last_loss = loss;
    int ctl = 0;
    if (it % 100 == 1) {
      if (fabs(min_loss - converge_min_loss) < 1e-5) ctl |= 1;
      else converge_min_loss = min_loss;
    }
    if (!(ctl & 1)) {
      if (loss < min_loss) {
        min_loss = loss;
        ctl |= 2;
        since_improve = 0;
      } else {
        ++since_improve;
      }
      if (since_improve >= 1000) ctl |= 1;
    }
    if (ctl & 1) stop_iter = it;
    return ctl;
  }
};
// Adam's bias corrections of iteration `it` (lr / (1 - b1^step) and sqrt(1 - b2^step)): a function of
// the iteration number alone, so another warp evaluates the two pow() while the loss is being formed.
__device__ __forceinline__ void wf_bias(int it, float* step_size, float* bc2_sqrt) {
  const int step = it + 1;
  *step_size = (float)((double)0.1 / (1.0 - pow((double)0.9, (double)step)));
  *bc2_sqrt = (float)sqrt(1.0 - pow((double)0.999, (double)step));
}
// Sum of the chunk sums in the one canonical order: lane l adds chunks l, l+32, l+64, ... in that
// order, then the 32 lane sums go through the warp shuffle tree.  Called by one full warp.
__device__ __forceinline__ double wf_sum_chunks(const volatile double* chunk, int64_t n_chunks, int lane) {
  double s = 0.0;
  for (int64_t c = lane; c < n_chunks; c += 32) s += chunk[c];
  return warp_sum(s);
}

// The loss is reduced in ONE order everywhere: e_t summed over chunks of 32 consecutive frames by a
// warp shuffle tree (frames past the end count as 0), then the chunk sums by wf_sum_chunks.

// AMP: every candidate row is scaled by amp[t,k] before mixing (compute_weight_with_amp,
// ddsp_prematch_dataset.py:684-803: `synth_set[...] * amp_ratio[:, :, None]` for all three
// neighbour offsets), i.e. the loss is the same quadratic form in a (.) w instead of w.
template <bool AMP>
__global__ void __launch_bounds__(WF_THREADS, 1) weight_fit_kernel(const double* __restrict__ gram_all,
                                                                   const int64_t* __restrict__ utt_offsets,
                                                                   int64_t n_pairs_total, int dim, double loss_scale,
                                                                   int max_iters, WfState st_all, int smem_floats,
                                                                   const float* __restrict__ amp_all,
                                                                   float* __restrict__ out_weights_all,
                                                                   double* __restrict__ info_all) {
  // one CTA per utterance: frames [f_begin, f_end) of the concatenated batch; every array is
  // indexed by the global frame number, so the CTA just offsets its bases
  extern __shared__ __align__(16) float s_w_dyn[];
  __shared__ double s_chunk[WF_SMEM_FRAMES / 32];
  __shared__ int s_ctl;  // bit0: stop, bit1: snapshot
  __shared__ float s_step_size, s_bc2_sqrt;
  const int64_t f_begin = utt_offsets[blockIdx.x], f_end = utt_offsets[blockIdx.x + 1];
  const int64_t T = f_end - f_begin;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* out_weights = out_weights_all + f_begin * 4;
  double* info = info_all ? info_all + (int64_t)blockIdx.x * 4 : nullptr;
  if (T <= 0) return;
  if (T < 2) {  // nothing to smooth: the reference's loop never improves on theta = 0 -> softmax = 1/4
    if (tid < 4) out_weights[tid] = 0.25f;
    if (tid == 0 && info) info[0] = info[1] = info[2] = info[3] = 0.0;
    return;
  }
  // Storage by utterance length (the arithmetic is the same in all three cases, so results do not
  // depend on it): weights + the whole optimiser state in shared memory (every phase then runs at
  // shared-memory latency instead of L2 round trips), weights only, or everything in global memory.
  const int64_t T4 = T * 4;
  const bool state_smem = 7 * T4 <= (int64_t)smem_floats;
  const bool use_smem = state_smem || T4 <= (int64_t)smem_floats;
  WfState st;
  st.theta = state_smem ? s_w_dyn + 1 * T4 : st_all.theta + f_begin * 4;
  st.m = state_smem ? s_w_dyn + 2 * T4 : st_all.m + f_begin * 4;
  st.v = state_smem ? s_w_dyn + 3 * T4 : st_all.v + f_begin * 4;
  st.vmax = state_smem ? s_w_dyn + 4 * T4 : st_all.vmax + f_begin * 4;
  st.best = state_smem ? s_w_dyn + 5 * T4 : st_all.best + f_begin * 4;
  st.grad = state_smem ? s_w_dyn + 6 * T4 : st_all.grad + f_begin * 4;
  st.wglob = st_all.wglob + f_begin * 4;
  const int64_t n_chunks = (T + 31) / 32;
  volatile double* chunk = n_chunks <= WF_SMEM_FRAMES / 32 ? s_chunk : st_all.chunk + f_begin / 32 + blockIdx.x;
  const double* gram = gram_all + f_begin;   // entry e of pair (t, t+1): gram[e * n_pairs_total + t]
  const float* amp = AMP ? amp_all + f_begin * 4 : nullptr;
  volatile float* wbuf = use_smem ? s_w_dyn : st.wglob;
  const double norm = loss_scale / ((double)(T - 1) * (double)dim);

  for (int64_t t = tid; t < T; t += WF_THREADS)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      st.theta[t * 4 + k] = 0.f;
      st.m[t * 4 + k] = 0.f;
      st.v[t * 4 + k] = 0.f;
      st.vmax[t * 4 + k] = 0.f;
      st.best[t * 4 + k] = 0.f;
    }
  WfControl ctrl(max_iters);
  __syncthreads();

  for (int it = 0; it < max_iters; ++it) {
    // ---- w = softmax(theta)
    for (int64_t t = tid; t < T; t += WF_THREADS) {
      float th[4], w[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) th[k] = st.theta[t * 4 + k];
      softmax4(th, w);
#pragma unroll
      for (int k = 0; k < 4; ++k) wbuf[t * 4 + k] = w[k];
    }
    if (tid == 32) wf_bias(it, &s_step_size, &s_bc2_sqrt);
    __syncthreads();
    // ---- loss and dL/dw from the Gram blocks (a warp holds 32 consecutive frames: one chunk)
    const int64_t NP = n_pairs_total;
    for (int64_t tb = 0; tb < T; tb += WF_THREADS) {
      const int64_t t = tb + tid;
      double e_t = 0.0;
      if (t < T) {
        double wt[4], wp[4] = {0.0, 0.0, 0.0, 0.0}, wn[4] = {0.0, 0.0, 0.0, 0.0}, at[4], g[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          at[k] = AMP ? (double)amp[t * 4 + k] : 1.0;
          wt[k] = (double)wbuf[t * 4 + k] * at[k];
        }
        const bool has_prev = t >= 1, has_next = t + 1 < T;
        if (has_prev)
#pragma unroll
          for (int k = 0; k < 4; ++k) wp[k] = (double)wbuf[(t - 1) * 4 + k] * (AMP ? (double)amp[(t - 1) * 4 + k] : 1.0);
        if (has_next)
#pragma unroll
          for (int k = 0; k < 4; ++k) wn[k] = (double)wbuf[(t + 1) * 4 + k] * (AMP ? (double)amp[(t + 1) * 4 + k] : 1.0);
        const double* Gp = gram + (t - 1);
        const double* Gc = gram + t;
        e_t = wf_frame(has_prev, has_next, wt, wp, wn, [&](int e) { return Gp[(int64_t)e * NP]; },
                       [&](int e) { return Gc[(int64_t)e * NP]; }, g);
        // gradient wrt w, cast to fp32 where the fp64 graph meets the fp32 softmax output
#pragma unroll
        for (int k = 0; k < 4; ++k) st.grad[t * 4 + k] = (float)(2.0 * norm * g[k] * at[k]);
      }
      e_t = warp_sum(e_t);
      if (lane == 0 && tb + warp * 32 < T) chunk[(tb >> 5) + warp] = e_t;
    }
    __syncthreads();
    if (warp == 0) {
      const double s = wf_sum_chunks(chunk, n_chunks, lane);
      if (lane == 0) s_ctl = ctrl.update(s * norm, it);
    }
    __syncthreads();
    const int ctl = s_ctl;
    if (ctl & 2)
      for (int64_t t = tid; t < T; t += WF_THREADS)
#pragma unroll
        for (int k = 0; k < 4; ++k) st.best[t * 4 + k] = st.theta[t * 4 + k];
    if (ctl & 1) break;
    // ---- Adam (amsgrad) step on the logits
    const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
    for (int64_t t = tid; t < T; t += WF_THREADS) {
      float w[4], gw[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        w[k] = wbuf[t * 4 + k];
        gw[k] = st.grad[t * 4 + k];
      }
      wf_adam4(st.theta + t * 4, st.m + t * 4, st.v + t * 4, st.vmax + t * 4, w, gw, step_size, bc2_sqrt);
    }
    __syncthreads();
  }
  __syncthreads();
  for (int64_t t = tid; t < T; t += WF_THREADS) {
    float th[4], w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) th[k] = st.best[t * 4 + k];
    softmax4(th, w);
#pragma unroll
    for (int k = 0; k < 4; ++k) out_weights[t * 4 + k] = w[k];
  }
  if (tid == 0 && info) {
    info[0] = (double)ctrl.stop_iter;
    info[1] = ctrl.min_loss;
    info[2] = ctrl.first_loss;
    info[3] = ctrl.last_loss;
  }
}

// ---- Cluster variant for a FEW LONG utterances (single-utterance conversion, BASELINE cfg 2): the
// one-CTA kernel walks a 3001-frame utterance at ~35 us per Adam iteration (six frames per thread,
// every Gram entry an L2 round trip).  Here a cluster of 8 CTAs splits the utterance into
// contiguous runs of 32-frame chunks; each CTA keeps its run's Gram blocks, weights and optimiser
// state in shared memory, exchanges the two boundary weights and its chunk sums with its peers
// through distributed shared memory, and every CTA runs the (deterministic) stop rule itself —
// two cluster barriers per iteration, nothing touches L2 inside the loop.  Same arithmetic, same
// reduction order as the one-CTA kernel: results are bit-identical.
constexpr int WC_C = 8;            // CTAs per cluster (portable maximum)
constexpr int WC_THREADS = 384;          // 12 warps = the 12 chunks of 32 frames a rank gets of a 3001-frame utterance: one pass per phase
constexpr int WC_MIN_FRAMES = 512; // every rank gets at least two chunks

__host__ __device__ inline int64_t wc_first_chunk(int64_t n_chunks, int rank) { return n_chunks * rank / WC_C; }

template <bool AMP>
__global__ void __cluster_dims__(WC_C, 1, 1) __launch_bounds__(WC_THREADS, 1)
weight_fit_cluster_kernel(const double* __restrict__ gram_all, const int64_t* __restrict__ utt_offsets,
                          int64_t n_pairs_total, int dim, double loss_scale, int max_iters, int gram_in_smem,
                          const float* __restrict__ amp_all, float* __restrict__ out_weights_all,
                          double* __restrict__ info_all) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) unsigned char wc_smem[];
  __shared__ int s_ctl;
  __shared__ float s_step_size, s_bc2_sqrt;
  const int rank = (int)cluster.block_rank();
  const int utt = blockIdx.x / WC_C;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t f_begin = utt_offsets[utt], f_end = utt_offsets[utt + 1];
  const int64_t T = f_end - f_begin;                 // >= WC_MIN_FRAMES (host-checked)
  const int64_t n_chunks = (T + 31) / 32;
  const int64_t c0 = wc_first_chunk(n_chunks, rank), c1 = wc_first_chunk(n_chunks, rank + 1);
  const int64_t f0 = c0 * 32, f1 = c1 * 32 < T ? c1 * 32 : T;   // this CTA's frames [f0, f1) of the utterance
  const int n_loc = (int)(f1 - f0);
  const int64_t p0 = f0 > 0 ? f0 - 1 : 0, p1 = f1 - 1 < T - 2 ? f1 - 1 : T - 2;   // Gram pairs p0..p1 (inclusive)
  const int np_loc = (int)(p1 - p0 + 1);
  // shared-memory carve-up (the same offsets in every CTA of the cluster, sized for the largest rank)
  const int max_loc = (int)(((n_chunks + WC_C - 1) / WC_C + 1) * 32);
  double* s_chunk = reinterpret_cast<double*>(wc_smem);                       // [n_chunks] all chunk sums of the utterance
  float* s_w = reinterpret_cast<float*>(s_chunk + n_chunks);                  // [(max_loc + 2) * 4] halo | local | halo
  float* s_theta = s_w + (size_t)(max_loc + 2) * 4;
  float* s_m = s_theta + (size_t)max_loc * 4;
  float* s_v = s_m + (size_t)max_loc * 4;
  float* s_vmax = s_v + (size_t)max_loc * 4;
  float* s_best = s_vmax + (size_t)max_loc * 4;
  float* s_grad = s_best + (size_t)max_loc * 4;
  double* s_gram = reinterpret_cast<double*>(s_grad + (size_t)max_loc * 4);  // [WF_E][np_loc] when gram_in_smem
  const double* gram = gram_all + f_begin;
  const float* amp = AMP ? amp_all + f_begin * 4 : nullptr;
  float* out_weights = out_weights_all + f_begin * 4;
  const double norm = loss_scale / ((double)(T - 1) * (double)dim);
  const int64_t NP = n_pairs_total;

  if (gram_in_smem)
    for (int e = 0; e < WF_E; ++e)
      for (int p = tid; p < np_loc; p += WC_THREADS) s_gram[e * np_loc + p] = gram[(int64_t)e * NP + p0 + p];
  for (int i = tid; i < n_loc * 4; i += WC_THREADS) {
    s_theta[i] = 0.f;
    s_m[i] = 0.f;
    s_v[i] = 0.f;
    s_vmax[i] = 0.f;
    s_best[i] = 0.f;
  }
  WfControl ctrl(max_iters);
  // peers' views: the left neighbour's right halo and the right neighbour's left halo
  float* left_halo = nullptr;
  float* right_halo = nullptr;
  if (rank > 0) {
    const int64_t lc0 = wc_first_chunk(n_chunks, rank - 1);
    const int left_n = (int)(f0 - lc0 * 32);
    left_halo = cluster.map_shared_rank(s_w, rank - 1) + (size_t)(left_n + 1) * 4;
  }
  if (rank + 1 < WC_C) right_halo = cluster.map_shared_rank(s_w, rank + 1);
  cluster.sync();   // every CTA's shared memory is live before anybody writes into it

  for (int it = 0; it < max_iters; ++it) {
    // ---- w = softmax(theta); boundary frames also go into the neighbours' halos
    for (int tl = tid; tl < n_loc; tl += WC_THREADS) {
      float th[4], w[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) th[k] = s_theta[tl * 4 + k];
      softmax4(th, w);
#pragma unroll
      for (int k = 0; k < 4; ++k) s_w[(tl + 1) * 4 + k] = w[k];
      if (tl == 0 && left_halo)
#pragma unroll
        for (int k = 0; k < 4; ++k) left_halo[k] = w[k];
      if (tl == n_loc - 1 && right_halo)
#pragma unroll
        for (int k = 0; k < 4; ++k) right_halo[k] = w[k];
    }
    if (tid == WC_THREADS - 32) wf_bias(it, &s_step_size, &s_bc2_sqrt);
    cluster.sync();
    // ---- loss and dL/dw; chunk sums go to every CTA of the cluster
    for (int tb = 0; tb < n_loc; tb += WC_THREADS) {
      const int tl = tb + tid;
      const int64_t t = f0 + tl;
      double e_t = 0.0;
      if (tl < n_loc) {
        double wt[4], wp[4] = {0.0, 0.0, 0.0, 0.0}, wn[4] = {0.0, 0.0, 0.0, 0.0}, at[4], g[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          at[k] = AMP ? (double)amp[t * 4 + k] : 1.0;
          wt[k] = (double)s_w[(tl + 1) * 4 + k] * at[k];
        }
        const bool has_prev = t >= 1, has_next = t + 1 < T;
        if (has_prev)
#pragma unroll
          for (int k = 0; k < 4; ++k) wp[k] = (double)s_w[tl * 4 + k] * (AMP ? (double)amp[(t - 1) * 4 + k] : 1.0);
        if (has_next)
#pragma unroll
          for (int k = 0; k < 4; ++k) wn[k] = (double)s_w[(tl + 2) * 4 + k] * (AMP ? (double)amp[(t + 1) * 4 + k] : 1.0);
        if (gram_in_smem) {
          const double* Gp = s_gram + (t - 1 - p0);
          const double* Gc = s_gram + (t - p0);
          e_t = wf_frame(has_prev, has_next, wt, wp, wn, [&](int e) { return Gp[e * np_loc]; },
                         [&](int e) { return Gc[e * np_loc]; }, g);
        } else {
          const double* Gp = gram + (t - 1);
          const double* Gc = gram + t;
          e_t = wf_frame(has_prev, has_next, wt, wp, wn, [&](int e) { return Gp[(int64_t)e * NP]; },
                         [&](int e) { return Gc[(int64_t)e * NP]; }, g);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) s_grad[tl * 4 + k] = (float)(2.0 * norm * g[k] * at[k]);
      }
      e_t = warp_sum(e_t);
      const int64_t ch = c0 + (tb >> 5) + warp;
      if (ch < c1 && lane < WC_C) cluster.map_shared_rank(s_chunk, lane)[ch] = e_t;   // lane r -> CTA r
    }
    cluster.sync();
    if (warp == 0) {
      const double s = wf_sum_chunks(s_chunk, n_chunks, lane);
      if (lane == 0) s_ctl = ctrl.update(s * norm, it);
    }
    __syncthreads();
    const int ctl = s_ctl;
    if (ctl & 2)
      for (int i = tid; i < n_loc * 4; i += WC_THREADS) s_best[i] = s_theta[i];
    if (ctl & 1) break;
    const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
    for (int tl = tid; tl < n_loc; tl += WC_THREADS) {
      float w[4], gw[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        w[k] = s_w[(tl + 1) * 4 + k];
        gw[k] = s_grad[tl * 4 + k];
      }
      wf_adam4(s_theta + tl * 4, s_m + tl * 4, s_v + tl * 4, s_vmax + tl * 4, w, gw, step_size, bc2_sqrt);
    }
    __syncthreads();
  }
  __syncthreads();
  for (int tl = tid; tl < n_loc; tl += WC_THREADS) {
    float th[4], w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) th[k] = s_best[tl * 4 + k];
    softmax4(th, w);
#pragma unroll
    for (int k = 0; k < 4; ++k) out_weights[(f0 + tl) * 4 + k] = w[k];
  }
  if (rank == 0 && tid == 0 && info_all) {
    double* info = info_all + (int64_t)utt * 4;
    info[0] = (double)ctrl.stop_iter;
    info[1] = ctrl.min_loss;
    info[2] = ctrl.first_loss;
    info[3] = ctrl.last_loss;
  }
  cluster.sync();   // nobody leaves while a peer could still address its shared memory
}

static size_t wc_smem_bytes(int64_t T, bool gram_in_smem) {
  const int64_t n_chunks = (T + 31) / 32;
  const int64_t max_loc = ((n_chunks + WC_C - 1) / WC_C + 1) * 32;
  size_t b = (size_t)n_chunks * sizeof(double) + (size_t)(max_loc + 2) * 16 + (size_t)max_loc * 16 * 6;
  if (gram_in_smem) b += (size_t)(max_loc + 1) * WF_E * sizeof(double);
  return b + 16;
}

size_t weight_fit_workspace_bytes(int64_t n_query, int k, int n_utt) {
  if (n_query < 1) return 256;
  size_t gram = (size_t)n_query * WF_E * sizeof(double);
  size_t state = (size_t)n_query * k * sizeof(float) * 7;  // theta, m, v, vmax, best, grad scratch, wglob
  size_t chunk = ((size_t)n_query / 32 + (size_t)n_utt + 2) * sizeof(double);
  size_t offs = ((size_t)(n_utt + 1) * sizeof(int64_t) + 255) / 256 * 256;
  return gram + state + chunk + offs + 1024;
}

// Batched launch: utterance u owns frames [utt_offsets_host[u], utt_offsets_host[u+1]) of the
// concatenated idx / out_weights; one CTA runs one utterance's whole optimisation, so a batch of
// utterances (BASELINE cfg 5) fills the SMs; a launch of few, long utterances gives each of them a
// cluster of 8 CTAs instead.  info: double[n_utt][4].
int launch_weight_fit(const int64_t* idx, const RowTable& synth, int dim,
                      const int64_t* utt_offsets_host, int n_utt, int k, double loss_scale, int max_iters,
                      const float* amp, float* out_weights, double* info, void* workspace, cudaStream_t stream) {
  KNN_CHECK_ARG(k == WF_K, -3, "weight_fit: k=%d, only k=%d is on the reference path", k, WF_K);
  if (n_utt == 0) return 0;
  const int64_t n_query = utt_offsets_host[n_utt];
  KNN_CHECK_ARG(utt_offsets_host[0] == 0, -1, "weight_fit: utterance offsets must start at 0");
  for (int u = 0; u < n_utt; ++u)
    KNN_CHECK_ARG(utt_offsets_host[u + 1] >= utt_offsets_host[u], -1, "weight_fit: utterance offsets must be non-decreasing");
  if (n_query == 0) return 0;
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
  double* gram = reinterpret_cast<double*>(ws);
  double* chunk = gram + (size_t)n_query * WF_E;
  float* f = reinterpret_cast<float*>(chunk + ((size_t)n_query / 32 + (size_t)n_utt + 2));
  WfState st;
  const size_t n4 = (size_t)n_query * 4;
  st.theta = f;
  st.m = f + n4;
  st.v = f + 2 * n4;
  st.vmax = f + 3 * n4;
  st.best = f + 4 * n4;
  st.grad = f + 5 * n4;
  st.wglob = f + 6 * n4;
  st.chunk = chunk;
  int64_t* d_off = reinterpret_cast<int64_t*>(f + 7 * n4);
  d_off = reinterpret_cast<int64_t*>((reinterpret_cast<uintptr_t>(d_off) + 255) & ~(uintptr_t)255);
  KNN_CUDA(cudaMemcpyAsync(d_off, utt_offsets_host, (size_t)(n_utt + 1) * sizeof(int64_t), cudaMemcpyHostToDevice,
                           stream));
  const int64_t n_pairs_total = n_query - 1;
  if (n_pairs_total > 0) {
    // pairs that straddle two utterances are computed too (same pool, in-range rows) and never read
    weight_gram_kernel<<<(unsigned)ceil_div64(n_pairs_total, WG_WARPS), WG_WARPS * 32, 0, stream>>>(idx, synth, dim,
                                                                                                 n_query, gram);
    KNN_LAUNCH_CHECK();
  }
  int64_t min_len = INT64_MAX, max_len = 0;
  for (int u = 0; u < n_utt; ++u) {
    const int64_t len = utt_offsets_host[u + 1] - utt_offsets_host[u];
    min_len = len < min_len ? len : min_len;
    max_len = len > max_len ? len : max_len;
  }
  const int64_t np_arg = n_pairs_total > 0 ? n_pairs_total : 1;
  {
    static PerDevice a0, a1, a2, a3;
    const int max_bytes = WF_STATE_FRAMES * 28 * (int)sizeof(float);   // 212.8 KB > 10240 * 16 B
    KNN_SMEM_ATTR(a0, weight_fit_kernel<false>, max_bytes);
    KNN_SMEM_ATTR(a1, weight_fit_kernel<true>, max_bytes);
    KNN_SMEM_ATTR(a2, weight_fit_cluster_kernel<false>, kWcMaxSmem);
    KNN_SMEM_ATTR(a3, weight_fit_cluster_kernel<true>, kWcMaxSmem);
  }
  // few, long utterances: a cluster per utterance (its state has to fit the cluster's shared memory,
  // and the device — or the SM partition this context runs in — has to be able to host a cluster)
  static PerDevice cluster_cache;   // per device: 0 = not asked yet, 1 = no, 2 = yes
  std::atomic<int>* cl_slot = cluster_cache.slot();
  int cluster_ok = cl_slot ? cl_slot->load(std::memory_order_relaxed) - 1 : -1;
  if (cluster_ok < 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(WC_C);
    cfg.blockDim = dim3(WC_THREADS);
    cfg.dynamicSmemBytes = kWcMaxSmem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = WC_C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n_clusters = 0;
    const cudaError_t e = cudaOccupancyMaxActiveClusters(&n_clusters, weight_fit_cluster_kernel<false>, &cfg);
    if (e != cudaSuccess) (void)cudaGetLastError();
    cluster_ok = (e == cudaSuccess && n_clusters >= 1) ? 1 : 0;
    if (cl_slot) cl_slot->store(cluster_ok + 1, std::memory_order_relaxed);
  }
  if (cluster_ok && opt_weight_fit_cluster() && n_utt * WC_C <= 144 && min_len >= WC_MIN_FRAMES &&
      wc_smem_bytes(max_len, false) <= (size_t)kWcMaxSmem) {
    const bool gram_in_smem = wc_smem_bytes(max_len, true) <= (size_t)kWcMaxSmem;
    const size_t smem = wc_smem_bytes(max_len, gram_in_smem);
    if (amp)
      weight_fit_cluster_kernel<true><<<n_utt * WC_C, WC_THREADS, smem, stream>>>(
          gram, d_off, np_arg, dim, loss_scale, max_iters, (int)gram_in_smem, amp, out_weights, info);
    else
      weight_fit_cluster_kernel<false><<<n_utt * WC_C, WC_THREADS, smem, stream>>>(
          gram, d_off, np_arg, dim, loss_scale, max_iters, (int)gram_in_smem, nullptr, out_weights, info);
    KNN_LAUNCH_CHECK();
    return 0;
  }
  // shared memory of the launch: the largest need over its utterances (each CTA picks its own mode)
  int64_t smem_floats = 4;
  for (int u = 0; u < n_utt; ++u) {
    const int64_t len = utt_offsets_host[u + 1] - utt_offsets_host[u];
    const int64_t need = len <= WF_STATE_FRAMES ? 28 * len : (len <= WF_SMEM_FRAMES ? 4 * len : 0);
    smem_floats = need > smem_floats ? need : smem_floats;
  }
  const size_t smem = (size_t)smem_floats * sizeof(float);
  if (amp)
    weight_fit_kernel<true><<<n_utt, WF_THREADS, smem, stream>>>(gram, d_off, np_arg, dim, loss_scale, max_iters, st,
                                                                 (int)smem_floats, amp, out_weights, info);
  else
    weight_fit_kernel<false><<<n_utt, WF_THREADS, smem, stream>>>(gram, d_off, np_arg, dim, loss_scale, max_iters, st,
                                                                  (int)smem_floats, nullptr, out_weights, info);
  KNN_LAUNCH_CHECK();
  return 0;
}

}  // namespace knnsvc
