// K3 gather + weighted mix, K4 f0 re-rank, K5 greedy concatenation-cost re-selection.
#include "common.cuh"
#include "kernels.cuh"

namespace knnsvc {

// ------------------------------------------------------------------ K3 gather_mix
// out[t,:] = sum_k w[t,k] * pool[idx[t,k],:]   (w == nullptr -> 1/K each)
// ddsp_prematch_dataset.py:1348,1358,1364 (features), :1435,1444,1446 (harmonics),
// ddsp_matcher.py:578.  HBM-bound: K row reads + one row write per query frame.
// One thread per 4 output columns (float4) when the row length allows it, K
// independent 16-byte loads in flight per thread.
template <bool VEC>
__global__ void __launch_bounds__(256) gather_mix_kernel(const float* __restrict__ pool, int64_t n_pool, int dim,
                                                         const int64_t* __restrict__ idx,
                                                         const float* __restrict__ weights, int64_t n_query, int k,
                                                         float* __restrict__ out) {
  const int per_row = VEC ? dim / 4 : dim;
  const int64_t total = n_query * per_row;
  const float uniform = 1.0f / (float)k;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = e / per_row;
    const int c = (int)(e % per_row);
    if (VEC) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j = 0; j < k; ++j) {
        int64_t r = __ldg(idx + t * k + j);
        r = r < 0 ? 0 : (r >= n_pool ? n_pool - 1 : r);
        const float w = weights ? __ldg(weights + t * k + j) : uniform;
        const float4 v = __ldg(reinterpret_cast<const float4*>(pool + r * dim) + c);
        acc.x = fmaf(w, v.x, acc.x);
        acc.y = fmaf(w, v.y, acc.y);
        acc.z = fmaf(w, v.z, acc.z);
        acc.w = fmaf(w, v.w, acc.w);
      }
      reinterpret_cast<float4*>(out + t * dim)[c] = acc;
    } else {
      float acc = 0.f;
      for (int j = 0; j < k; ++j) {
        int64_t r = __ldg(idx + t * k + j);
        r = r < 0 ? 0 : (r >= n_pool ? n_pool - 1 : r);
        const float w = weights ? __ldg(weights + t * k + j) : uniform;
        acc = fmaf(w, __ldg(pool + r * dim + c), acc);
      }
      out[t * dim + c] = acc;
    }
  }
}

int launch_gather_mix(const float* pool, int64_t n_pool, int dim, const int64_t* idx, const float* weights,
                      int64_t n_query, int k, float* out, cudaStream_t stream) {
  if (n_query == 0 || dim == 0) return 0;
  const bool vec = (dim % 4 == 0) && ((reinterpret_cast<uintptr_t>(pool) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  const int64_t total = n_query * (vec ? dim / 4 : dim);
  int64_t grid = ceil_div64(total, 256);
  if (grid > 148 * 32) grid = 148 * 32;
  if (vec)
    gather_mix_kernel<true><<<(unsigned)grid, 256, 0, stream>>>(pool, n_pool, dim, idx, weights, n_query, k, out);
  else
    gather_mix_kernel<false><<<(unsigned)grid, 256, 0, stream>>>(pool, n_pool, dim, idx, weights, n_query, k, out);
  KNN_LAUNCH_CHECK();
  return 0;
}

// K3 over a pool that lives in several row blocks (sharded.py): the same arithmetic in the same
// order as gather_mix_kernel, so the result is bit-identical to mixing from one contiguous pool;
// rows of another GPU's shard are read through its IPC-mapped pointer (NVLink P2P loads, 16 B per
// lane, K of them in flight per thread).  Launched by the rank that OWNS the query rows.
template <bool VEC>
__global__ void __launch_bounds__(256) gather_mix_sharded_kernel(const __grid_constant__ RowTable tab, int dim,
                                                                 const int64_t* __restrict__ idx,
                                                                 const float* __restrict__ weights, int64_t n_query,
                                                                 int k, float* __restrict__ out) {
  const int per_row = VEC ? dim / 4 : dim;
  const int64_t total = n_query * per_row;
  const float uniform = 1.0f / (float)k;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = e / per_row;
    const int c = (int)(e % per_row);
    if (VEC) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j = 0; j < k; ++j) {
        const float* row = table_row(tab, __ldg(idx + t * k + j), dim);
        const float w = weights ? __ldg(weights + t * k + j) : uniform;
        const float4 v = __ldg(reinterpret_cast<const float4*>(row) + c);
        acc.x = fmaf(w, v.x, acc.x);
        acc.y = fmaf(w, v.y, acc.y);
        acc.z = fmaf(w, v.z, acc.z);
        acc.w = fmaf(w, v.w, acc.w);
      }
      reinterpret_cast<float4*>(out + t * dim)[c] = acc;
    } else {
      float acc = 0.f;
      for (int j = 0; j < k; ++j) {
        const float* row = table_row(tab, __ldg(idx + t * k + j), dim);
        const float w = weights ? __ldg(weights + t * k + j) : uniform;
        acc = fmaf(w, __ldg(row + c), acc);
      }
      out[t * dim + c] = acc;
    }
  }
}

int launch_gather_mix_sharded(const RowTable& tab, int dim, const int64_t* idx, const float* weights, int64_t n_query,
                              int k, float* out, cudaStream_t stream) {
  if (n_query == 0 || dim == 0) return 0;
  bool vec = (dim % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  for (int s = 0; s < tab.n; ++s) vec = vec && ((reinterpret_cast<uintptr_t>(tab.base[s]) & 15) == 0);
  const int64_t total = n_query * (vec ? dim / 4 : dim);
  int64_t grid = ceil_div64(total, 256);
  if (grid > 148 * 32) grid = 148 * 32;
  if (vec)
    gather_mix_sharded_kernel<true><<<(unsigned)grid, 256, 0, stream>>>(tab, dim, idx, weights, n_query, k, out);
  else
    gather_mix_sharded_kernel<false><<<(unsigned)grid, 256, 0, stream>>>(tab, dim, idx, weights, n_query, k, out);
  KNN_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------ K4 f0 re-rank
// sort_by_f0_compatibility (ddsp_prematch_dataset.py:954-997): key =
// |log2(f0[idx]+1e-5) - log2(expected+1e-5)| in fp32, stable ascending sort of the
// k <= 32 candidates of a row.  One warp per row, one candidate per lane; the
// stable rank is a count over the warp (a 32-wide sorting network by shuffles).
__global__ void __launch_bounds__(256) f0_rerank_kernel(const float* __restrict__ expected_f0,
                                                        const float* __restrict__ pool_f0,
                                                        const int64_t* __restrict__ idx, int64_t n_query, int k,
                                                        int64_t* __restrict__ out_idx) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t t = warp; t < n_query; t += nwarps) {
    const float e = log2f(__ldg(expected_f0 + t) + 1e-5f);
    int64_t my = -1;
    float key = INFINITY;
    if (lane < k) {
      my = __ldg(idx + t * k + lane);
      key = fabsf(log2f(__ldg(pool_f0 + my) + 1e-5f) - e);
    }
    int rank = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float kj = __shfl_sync(0xffffffffu, key, j);
      rank += (j < k) && ((kj < key) || (kj == key && j < lane));
    }
    if (lane < k) out_idx[t * k + rank] = my;
  }
}

int launch_f0_rerank(const float* expected_f0, const float* pool_f0, const int64_t* idx, int64_t n_query, int k,
                     int64_t* out_idx, cudaStream_t stream) {
  if (n_query == 0) return 0;
  KNN_CHECK_ARG(k >= 1 && k <= 32, -3, "f0_rerank: k=%d outside [1,32]", k);
  int64_t grid = ceil_div64(n_query, 8);
  if (grid > 148 * 16) grid = 148 * 16;
  f0_rerank_kernel<<<(unsigned)grid, 256, 0, stream>>>(expected_f0, pool_f0, idx, n_query, k, out_idx);
  KNN_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------ K5 greedy re-selection
// knn_with_concat_cost (lib_ongaku_test.py:270-369), K = 4.  This is the GENERAL kernel (any
// feature dimension / alignment, rows read through L2 every step); feature rows of up to
// 1024 aligned floats take the shared-memory staged kernel in concat_cost_sm100.cu.  The recurrence is
// serial in the frame index (frame i's candidates include "previous selection + 1"),
// so one CTA walks one utterance; parallelism is inside a step — 8 warps, one
// candidate row each, 5 direct-form cosine distances per candidate — and across
// utterances (one CTA each).  Everything that does NOT depend on the recurrence
// is hoisted into a fully parallel pre-pass: the frame-to-frame baseline
// 2*dist(src[i-1], src[i]) (:310) and |src[i]|^2.  Norms of the previous
// selections are carried from the step in which they were candidates.
//
// Distances follow the reference's small-matrix path: cdist evaluates
// sum((x-y)^2) directly, dot = (-d2 + |x|^2 + |y|^2)/2,
// dist = 1 - dot/(|x||y|).  Differences and squares are fp32 (the operands are
// fp32), four-term partials are summed in fp64.
constexpr int CC_K = 4;
constexpr int CC_C = 2 * CC_K;

__device__ __forceinline__ double cosd_from(double d2, double nx2, double ny2) {
  const double dot = (-d2 + nx2 + ny2) * 0.5;
  return 1.0 - dot / (sqrt(nx2) * sqrt(ny2));
}

// one warp per frame: base[i] = 2*cosd(src[i-1], src[i]) (i >= 1), n2[i] = |src[i]|^2
__global__ void __launch_bounds__(256) frame_baseline_kernel(const float* __restrict__ src, int dim, int64_t n_frames,
                                                             double* __restrict__ base, double* __restrict__ n2) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t i = warp; i < n_frames; i += nwarps) {
    const float* b = src + i * dim;
    const float* a = src + (i > 0 ? i - 1 : 0) * dim;
    double na = 0, nb = 0, dd = 0;
    for (int c = lane; c < dim; c += 32) {
      const float av = __ldg(a + c), bv = __ldg(b + c);
      na += (double)(av * av);
      nb += (double)(bv * bv);
      dd += (double)((av - bv) * (av - bv));
    }
    na = warp_sum(na); nb = warp_sum(nb); dd = warp_sum(dd);
    if (lane == 0) {
      n2[i] = nb;
      base[i] = (i > 0) ? 2.0 * cosd_from(dd, na, nb) : 0.0;
    }
  }
}

// words of a small device buffer -> mapped pinned host memory, by stores (no copy engine involved)
__global__ void __launch_bounds__(256) store_to_host_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst,
                                                            size_t n_words) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
}
int launch_store_to_host(const void* src, void* dst_mapped, size_t nbytes, cudaStream_t stream) {
  const size_t n_words = nbytes / 4;
  size_t grid = (n_words + 255) / 256;
  if (grid > 64) grid = 64;      // a few SMs saturate PCIe
  store_to_host_kernel<<<(unsigned)grid, 256, 0, stream>>>(reinterpret_cast<const uint32_t*>(src),
                                                          reinterpret_cast<uint32_t*>(dst_mapped), n_words);
  KNN_LAUNCH_CHECK();
  return 0;
}

// log2(f0 + 1e-5) of every pool frame, once per call: the cluster kernel's warps then load a candidate's value
// instead of running log2 next to the recurrence (lib_ongaku_test.py:322-323 takes the log of both f0 tracks)
__global__ void __launch_bounds__(256) log_f0_table_kernel(const float* __restrict__ f0, int64_t n, double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = log2((double)__ldg(f0 + i) + 1e-5);
}

// 16 warps: warp w scores candidate (w >> 1) over half (w & 1) of the feature dimension, so one
// batch of 6 rows x 4 float4 loads per lane covers a 1024-dim row and a step costs about one
// L2 round trip plus ~600 instructions per warp.
constexpr int CC_WARPS = 2 * CC_C;
constexpr int CC_ACC = 6;  // |c|^2, d2(src,c), d2(prev_j,c) j=0..3

__global__ void __launch_bounds__(CC_WARPS * 32) concat_cost_kernel(
    const int64_t* __restrict__ idx, const float* __restrict__ src, const __grid_constant__ RowTable pool,
    int dim, const float* __restrict__ src_f0, const float* __restrict__ pool_f0, float concat_weight,
    const int64_t* __restrict__ utt_offsets, const double* __restrict__ base_all, const double* __restrict__ src_n2,
    int64_t* __restrict__ out_idx) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t f_begin = utt_offsets[blockIdx.x], f_end = utt_offsets[blockIdx.x + 1];
  if (f_end <= f_begin) return;
  const bool use_f0 = src_f0 != nullptr;
  const bool vec4 = (dim % 8 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && table_aligned16(pool);
  const int64_t n_pool = pool.lo[pool.n];

  __shared__ int64_t s_prev[CC_K];
  __shared__ double s_prev_n2[CC_K];
  __shared__ int64_t s_cand[CC_C];
  __shared__ double s_part[CC_WARPS][CC_ACC];
  __shared__ double s_w;

  if (threadIdx.x < CC_K) {
    const int64_t v = idx[f_begin * CC_K + threadIdx.x];
    s_prev[threadIdx.x] = v;
    out_idx[f_begin * CC_K + threadIdx.x] = v;
  }
  if (threadIdx.x == 0) s_w = (double)concat_weight;
  __syncthreads();
  if (warp < CC_K) {  // |row|^2 of the four initial selections
    const float* r = table_row(pool, s_prev[warp], dim);
    double n = 0;
    for (int c = lane; c < dim; c += 32) {
      const float v = __ldg(r + c);
      n += (double)(v * v);
    }
    n = warp_sum(n);
    if (lane == 0) s_prev_n2[warp] = n;
  }
  // values that do not depend on the recurrence are fetched one step ahead
  int64_t next_idx = 0;
  double next_base = 0.0, next_n2 = 0.0, next_lsrc = 0.0;
  if (f_begin + 1 < f_end) {
    if (threadIdx.x < CC_K) next_idx = idx[(f_begin + 1) * CC_K + threadIdx.x];
    if (threadIdx.x < CC_C) {
      next_base = base_all[f_begin + 1];
      next_n2 = src_n2[f_begin + 1];
      if (use_f0) next_lsrc = log2((double)__ldg(src_f0 + f_begin + 1) + 1e-5);
    }
  }
  __syncthreads();

  const int cand_id = warp >> 1, half = warp & 1;
  for (int64_t i = f_begin + 1; i < f_end; ++i) {
    const double base = next_base, sn2 = next_n2, lsrc = next_lsrc;   // meaningful on threads < CC_C
    if (threadIdx.x < CC_C) {
      int64_t c;
      if (threadIdx.x < CC_K) {
        c = next_idx;
      } else {
        c = s_prev[threadIdx.x - CC_K] + 1;
        if (c >= n_pool) c = n_pool - 1;
      }
      s_cand[threadIdx.x] = c;
    }
    __syncthreads();
    if (i + 1 < f_end) {  // issue next step's independent loads now; they land while this step computes
      if (threadIdx.x < CC_K) next_idx = idx[(i + 1) * CC_K + threadIdx.x];
      if (threadIdx.x < CC_C) {
        next_base = base_all[i + 1];
        next_n2 = src_n2[i + 1];
        if (use_f0) next_lsrc = log2((double)__ldg(src_f0 + i + 1) + 1e-5);
      }
    }
    double lcand = 0.0;
    if (use_f0 && threadIdx.x < CC_C) lcand = log2((double)__ldg(pool_f0 + s_cand[threadIdx.x]) + 1e-5);
    {
      const float* crow = table_row(pool, s_cand[cand_id], dim);
      const float* srow = src + i * dim;
      const float* p0 = table_row(pool, s_prev[0], dim);
      const float* p1 = table_row(pool, s_prev[1], dim);
      const float* p2 = table_row(pool, s_prev[2], dim);
      const float* p3 = table_row(pool, s_prev[3], dim);
      double nc = 0, dm = 0, d0 = 0, d1 = 0, d2 = 0, d3 = 0;
      if (vec4) {
        const int n4h = dim / 8;                 // float4s in this warp's half of the row
        const int off = half * n4h;
        const float4* c4 = reinterpret_cast<const float4*>(crow) + off;
        const float4* s4 = reinterpret_cast<const float4*>(srow) + off;
        const float4* q0 = reinterpret_cast<const float4*>(p0) + off;
        const float4* q1 = reinterpret_cast<const float4*>(p1) + off;
        const float4* q2 = reinterpret_cast<const float4*>(p2) + off;
        const float4* q3 = reinterpret_cast<const float4*>(p3) + off;
        constexpr int U = 4;  // 6 rows x U float4 loads issued before any use
        for (int c0 = lane; c0 < n4h; c0 += 32 * U) {
          float4 cv[U], sv[U], a0[U], a1[U], a2[U], a3[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int c = c0 + 32 * u;
            const bool ok = c < n4h;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            cv[u] = ok ? __ldg(c4 + c) : z;
            sv[u] = ok ? __ldg(s4 + c) : z;
            a0[u] = ok ? __ldg(q0 + c) : z;
            a1[u] = ok ? __ldg(q1 + c) : z;
            a2[u] = ok ? __ldg(q2 + c) : z;
            a3[u] = ok ? __ldg(q3 + c) : z;
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            float e;
            float t_nc = cv[u].x * cv[u].x;
            t_nc = fmaf(cv[u].y, cv[u].y, t_nc);
            t_nc = fmaf(cv[u].z, cv[u].z, t_nc);
            t_nc = fmaf(cv[u].w, cv[u].w, t_nc);
#define KNN_D2(acc, a)                                                  \
            {                                                           \
              float t;                                                  \
              e = a.x - cv[u].x; t = e * e;                             \
              e = a.y - cv[u].y; t = fmaf(e, e, t);                     \
              e = a.z - cv[u].z; t = fmaf(e, e, t);                     \
              e = a.w - cv[u].w; t = fmaf(e, e, t);                     \
              acc += (double)t;                                         \
            }
            nc += (double)t_nc;
            KNN_D2(dm, sv[u]) KNN_D2(d0, a0[u]) KNN_D2(d1, a1[u]) KNN_D2(d2, a2[u]) KNN_D2(d3, a3[u])
#undef KNN_D2
          }
        }
      } else {
        const int h0 = half * ((dim + 1) / 2), h1 = half ? dim : (dim + 1) / 2;
        for (int c = h0 + lane; c < h1; c += 32) {
          const float cv = __ldg(crow + c), sv = __ldg(srow + c);
          const float a0 = __ldg(p0 + c), a1 = __ldg(p1 + c), a2 = __ldg(p2 + c), a3 = __ldg(p3 + c);
          nc += (double)(cv * cv);
          dm += (double)((sv - cv) * (sv - cv));
          d0 += (double)((a0 - cv) * (a0 - cv));
          d1 += (double)((a1 - cv) * (a1 - cv));
          d2 += (double)((a2 - cv) * (a2 - cv));
          d3 += (double)((a3 - cv) * (a3 - cv));
        }
      }
      nc = warp_sum(nc); dm = warp_sum(dm);
      d0 = warp_sum(d0); d1 = warp_sum(d1); d2 = warp_sum(d2); d3 = warp_sum(d3);
      if (lane == 0) {
        s_part[warp][0] = nc; s_part[warp][1] = dm;
        s_part[warp][2] = d0; s_part[warp][3] = d1; s_part[warp][4] = d2; s_part[warp][5] = d3;
      }
    }
    __syncthreads();
    if (warp == 0) {
      double w = s_w;
      if (use_f0 && !(base < 0.08)) w = 0.0;  // sticky: persists for all later frames (lib_ongaku_test.py:332)
      double total = INFINITY, my_n2 = 0.0;
      if (lane < CC_C) {
        double v[CC_ACC];
#pragma unroll
        for (int a = 0; a < CC_ACC; ++a) v[a] = s_part[2 * lane][a] + s_part[2 * lane + 1][a];
        my_n2 = v[0];
        const double match = cosd_from(v[1], sn2, my_n2);
        double cc[CC_K];
#pragma unroll
        for (int j = 0; j < CC_K; ++j) cc[j] = cosd_from(v[2 + j], s_prev_n2[j], my_n2);
        if (use_f0) {
          if (base < 0.08) {
#pragma unroll
            for (int j = 0; j < CC_K; ++j)
              if (cc[j] < 5.0 * base) cc[j] = 0.0;
          }
        } else {
#pragma unroll
          for (int j = 0; j < CC_K; ++j)
            if (cc[j] > base) cc[j] = 1.5 * cc[j] - base;
        }
        // lower median of 4 = second smallest (torch.median, lib_ongaku_test.py:337,342)
        double lo01 = fmin(cc[0], cc[1]), hi01 = fmax(cc[0], cc[1]);
        double lo23 = fmin(cc[2], cc[3]), hi23 = fmax(cc[2], cc[3]);
        double med = fmin(fmax(lo01, lo23), fmin(hi01, hi23));
        total = w * med + match;
        if (use_f0) total += fabs(lcand - lsrc);
      }
      int rank = 0;
#pragma unroll
      for (int j = 0; j < CC_C; ++j) {
        const double tj = __shfl_sync(0xffffffffu, total, j);
        rank += (tj < total) || (tj == total && j < lane);
      }
      __syncwarp();  // every lane has read s_prev_n2 before it is overwritten
      if (lane < CC_C && rank < CC_K) {
        const int64_t sel = s_cand[lane];
        out_idx[i * CC_K + rank] = sel;
        s_prev[rank] = sel;
        s_prev_n2[rank] = my_n2;
      }
      if (lane == 0) s_w = w;
    }
    __syncthreads();
  }
}

int launch_concat_cost(const int64_t* idx, const float* src, const RowTable& pool, int dim,
                       const float* src_f0, const float* pool_f0, float concat_weight, const int64_t* utt_offsets_dev,
                       int n_utt, int64_t n_frames, double* frame_ws, double* lf0_ws, int64_t* out_idx,
                       cudaStream_t stream) {
  if (n_utt == 0 || n_frames == 0) return 0;
  double* base = frame_ws;
  double* n2 = frame_ws + n_frames;
  int64_t grid = ceil_div64(n_frames, 8);
  if (grid > 148 * 16) grid = 148 * 16;
  frame_baseline_kernel<<<(unsigned)grid, 256, 0, stream>>>(src, dim, n_frames, base, n2);
  KNN_LAUNCH_CHECK();
  if (opt_concat_staged() && concat_staged_eligible(src, pool, dim)) {   // shared-memory staged recurrence (concat_cost_sm100.cu)
    if (opt_concat_cluster() && concat_cluster_fits(n_utt) && n_frames < ((int64_t)1 << 31)) {   // few utterances: 8 SMs each (32-bit frame counters)
      if (src_f0 && lf0_ws) {
        const int64_t n_pool = pool.lo[pool.n];
        int64_t g2 = ceil_div64(n_pool, 256);
        if (g2 > 148 * 8) g2 = 148 * 8;
        log_f0_table_kernel<<<(unsigned)g2, 256, 0, stream>>>(pool_f0, n_pool, lf0_ws);
        KNN_LAUNCH_CHECK();
      }
      return launch_concat_cost_cluster(idx, src, pool, dim, src_f0, pool_f0, src_f0 ? lf0_ws : nullptr, concat_weight,
                                        utt_offsets_dev, n_utt, base, n2, out_idx, stream);
    }
    return launch_concat_cost_staged(idx, src, pool, dim, src_f0, pool_f0, concat_weight, utt_offsets_dev, n_utt,
                                     base, n2, out_idx, stream);
  }
  concat_cost_kernel<<<n_utt, CC_WARPS * 32, 0, stream>>>(idx, src, pool, dim, src_f0, pool_f0, concat_weight,
                                                      utt_offsets_dev, base, n2, out_idx);
  KNN_LAUNCH_CHECK();
  return 0;
}

}  // namespace knnsvc
