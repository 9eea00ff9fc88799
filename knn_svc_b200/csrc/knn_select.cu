// Exact re-scoring of the tensor-core filter's candidate log, and the top-k
// merge used after the NCCL all-gather of per-shard results (C1).
//
// knn_rescore: one CTA per query row.
//   1. tau* = k-th largest approximate similarity over the row's per-segment
//      top-k lists; every true top-k member has s~ >= tau* - 2*eps.
//   2. survivors = logged candidates with s~ >= tau* - 2*eps.
//   3. each survivor is re-scored from the fp32 rows with fp64 accumulation:
//      dist = 1 - q.p/(|q||p|)   (lib_ongaku_test.py:162-165).
//   4. survivors are ranked by (dist, index) and the first k written, ascending —
//      the order dists.topk(k, largest=False) returns (ddsp_prematch_dataset.py:1203).
// Survivors are compacted into a shared-memory list (any number: the list is reduced to its
// best k whenever it fills), scored two per warp pass, and ranked by counting.  Rows whose log
// overflowed in some segment (more than `cap` candidates inside the window:
// massive ties) are appended to the flag list for the exact brute-force kernel.
#include <limits.h>

#include "common.cuh"
#include "kernels.cuh"

namespace knnsvc {

constexpr int RS_THREADS = 128;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_MAXTOP = 16 * kMaxK;  // n_seg <= 16
constexpr int RS_LIST = 1024;          // survivors held in shared memory between two reductions to the best k

__device__ __forceinline__ bool rs_less(double da, int ia, double db, int ib) {
  return da < db || (da == db && ia < ib);
}

// dot products of ONE or TWO pool rows with the query row held in shared memory as fp64
// (products of two fp32 values are exact in fp64; the accumulation order is fixed, so every
// kernel that scores a pair gets the same bits).  Two rows per pass share the query loads and
// give the scheduler two independent DFMA chains.
template <bool VEC>
__device__ __forceinline__ void rs_dot2(const float* __restrict__ pa, const float* __restrict__ pb,
                                        const double* __restrict__ s_q, int dim, int lane, double& ra, double& rb) {
  double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
  if (VEC) {
    const float4* a4 = reinterpret_cast<const float4*>(pa);
    const float4* b4 = reinterpret_cast<const float4*>(pb);
    // VEC layout of the query: element pairs (4c, 4c+1) in the first half, (4c+2, 4c+3) in the
    // second, so consecutive lanes read consecutive 16-byte words (no bank conflicts)
    const double2* q2 = reinterpret_cast<const double2*>(s_q);
    const int nv = dim / 4;
#pragma unroll 4
    for (int c = lane; c < nv; c += 32) {
      const float4 va = __ldg(a4 + c);
      const float4 vb = __ldg(b4 + c);
      const double2 qa = q2[c], qb = q2[nv + c];
      a0 = fma((double)va.x, qa.x, a0);
      a1 = fma((double)va.y, qa.y, a1);
      a0 = fma((double)va.z, qb.x, a0);
      a1 = fma((double)va.w, qb.y, a1);
      b0 = fma((double)vb.x, qa.x, b0);
      b1 = fma((double)vb.y, qa.y, b1);
      b0 = fma((double)vb.z, qb.x, b0);
      b1 = fma((double)vb.w, qb.y, b1);
    }
  } else {
    for (int c = lane; c < dim; c += 32) {
      const double qv = s_q[c];
      a0 = fma((double)__ldg(pa + c), qv, a0);
      b0 = fma((double)__ldg(pb + c), qv, b0);
    }
  }
  ra = a0 + a1;
  rb = b0 + b1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ra += __shfl_xor_sync(0xffffffffu, ra, o);
    rb += __shfl_xor_sync(0xffffffffu, rb, o);
  }
}

__global__ void __launch_bounds__(RS_THREADS) knn_rescore_kernel(
    const float* __restrict__ q, const double* __restrict__ qn, int64_t n_query, const float* __restrict__ p,
    const double* __restrict__ pn, int64_t n_pool, int dim, int k, int n_seg, int cap,
    const float* __restrict__ log_val, const int* __restrict__ log_idx, const int* __restrict__ log_cnt,
    const float* __restrict__ seg_top, int64_t index_offset, float* __restrict__ out_dist,
    double* __restrict__ out_dist64, int64_t* __restrict__ out_idx, int64_t* __restrict__ flag_list,
    int* __restrict__ flag_count,
    int* __restrict__ stats, const int64_t* __restrict__ mask_lo, const int64_t* __restrict__ mask_hi,
    const float* __restrict__ q_err, const float* __restrict__ p_err) {
  extern __shared__ __align__(16) double s_q[];  // [dim] the query row, converted once
  __shared__ float s_top[RS_MAXTOP];
  __shared__ double s_dist[RS_LIST];             // exact distances of the survivors in s_cand
  __shared__ int s_cand[RS_LIST];
  __shared__ double s_best_d[kMaxK];             // scratch for the reduction to the best k
  __shared__ int s_best_i[kMaxK];
  __shared__ int s_n, s_overflow, s_logged, s_nsurv;
  __shared__ float s_thr;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool vec = (dim & 3) == 0 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0);

  for (int64_t row = blockIdx.x; row < n_query; row += gridDim.x) {
    __syncthreads();
    if (tid == 0) {
      s_overflow = 0;
      s_logged = 0;
      s_nsurv = 0;
      s_n = 0;
      s_thr = -INFINITY;
    }
    const int n_top = n_seg * k;
    for (int e = tid; e < n_top; e += RS_THREADS) s_top[e] = seg_top[row * n_top + e];
    if (vec) {
      const int nv = dim / 4;
      for (int c = tid; c < dim; c += RS_THREADS) {
        const int g4 = c >> 2, r = c & 3;
        s_q[(r < 2 ? 0 : 2 * nv) + 2 * g4 + (r & 1)] = (double)__ldg(q + row * dim + c);
      }
    } else {
      for (int c = tid; c < dim; c += RS_THREADS) s_q[c] = (double)__ldg(q + row * dim + c);
    }
    __syncthreads();
    // k-th largest of the union of the segment lists (rank counting, ties by position)
    for (int e = tid; e < n_top; e += RS_THREADS) {
      const float v = s_top[e];
      int rank = 0;
      for (int j = 0; j < n_top; ++j) {
        const float o = s_top[j];
        rank += (o > v) || (o == v && j < e);
      }
      if (rank == k - 1) s_thr = v - 2.0f * filter_eps(q_err, p_err, (dim + 63) / 64 * 64);
    }
    for (int s = tid; s < n_seg; s += RS_THREADS) {
      const int c = log_cnt[row * n_seg + s];
      if (c > cap) s_overflow = 1;
      atomicAdd(&s_logged, c > cap ? cap : c);
    }
    __syncthreads();
    if (s_overflow) {
      // more than `cap` candidates inside the error window of one segment: exact brute force decides
      if (tid == 0) {
        const int pos = atomicAdd(flag_count, 1);
        flag_list[pos] = row;
        if (stats) atomicAdd(stats + 0, 1);
      }
      continue;
    }
    const float thr = s_thr;
    const double qnorm = qn[row];
    // masked column range: distance defined as 1 (ddsp_prematch_dataset.py:1623-1624)
    const int64_t m_lo = mask_lo ? mask_lo[row] : 0, m_hi = mask_lo ? mask_hi[row] : 0;
    int n_scored = 0;   // entries [0, n_scored) of the list already carry their exact distance

    // score list entries [n_scored, s_n): a warp takes two survivors per pass
    auto score = [&]() {
      const int n = s_n;
      for (int e = n_scored + 2 * warp; e < n; e += 2 * RS_WARPS) {
        const int ia = s_cand[e];
        const bool has_b = e + 1 < n;
        const int ib = has_b ? s_cand[e + 1] : ia;
        double da, db;
        if (vec) rs_dot2<true>(p + (int64_t)ia * dim, p + (int64_t)ib * dim, s_q, dim, lane, da, db);
        else rs_dot2<false>(p + (int64_t)ia * dim, p + (int64_t)ib * dim, s_q, dim, lane, da, db);
        if (lane == 0) {
          s_dist[e] = (ia >= m_lo && ia < m_hi) ? 1.0 : 1.0 - da / (qnorm * pn[ia]);
          if (has_b) s_dist[e + 1] = (ib >= m_lo && ib < m_hi) ? 1.0 : 1.0 - db / (qnorm * pn[ib]);
        }
      }
      n_scored = n;
    };
    // keep the best k of the scored list (rank by (dist, index)); final: write them out in order
    auto reduce_to_best = [&](bool final) {
      const int n = s_n;
      for (int e = tid; e < n; e += RS_THREADS) {
        const double d = s_dist[e];
        const int i = s_cand[e];
        int rank = 0;
        for (int j = 0; j < n; ++j) rank += rs_less(s_dist[j], s_cand[j], d, i);
        if (rank < k) {
          if (final) {
            out_dist[row * k + rank] = (float)d;
            if (out_dist64) out_dist64[row * k + rank] = d;
            out_idx[row * k + rank] = (int64_t)i + index_offset;
          } else {
            s_best_d[rank] = d;
            s_best_i[rank] = i;
          }
        }
      }
      if (!final) {
        __syncthreads();
        const int keep = n < k ? n : k;
        if (tid < keep) {
          s_dist[tid] = s_best_d[tid];
          s_cand[tid] = s_best_i[tid];
        }
        if (tid == 0) s_n = keep;
        n_scored = keep;
        __syncthreads();
      }
    };

    for (int s = 0; s < n_seg; ++s) {
      const int64_t slot = row * n_seg + s;
      const int c = log_cnt[slot];
      for (int base = 0; base < c; base += RS_THREADS) {
        if (s_n + RS_THREADS > RS_LIST) {   // block-uniform (s_n is read after a barrier)
          score();
          __syncthreads();
          reduce_to_best(false);
        }
        const int e = base + tid;
        if (e < c && log_val[slot * cap + e] >= thr) {
          const int pos = atomicAdd(&s_n, 1);
          s_cand[pos] = log_idx[slot * cap + e];
          atomicAdd(&s_nsurv, 1);
        }
        __syncthreads();
      }
    }
    score();
    __syncthreads();
    reduce_to_best(true);
    if (tid == 0 && stats) {
      atomicAdd(stats + 1, s_logged);
      atomicAdd(stats + 2, s_nsurv);
    }
  }
}

int launch_knn_rescore(const float* q, const double* qn, int64_t n_query, const float* p, const double* pn,
                       int64_t n_pool, int dim, int k, const FilterPlan& pl, const float* log_val,
                       const int* log_idx, const int* log_cnt, const float* seg_top, int64_t index_offset,
                       float* out_dist, double* out_dist64, int64_t* out_idx, int64_t* flag_list, int* flag_count,
                       int* stats, const int64_t* mask_lo, const int64_t* mask_hi, const float* q_err,
                       const float* p_err, cudaStream_t stream) {
  if (n_query == 0) return 0;
  KNN_CHECK_ARG(pl.n_seg * k <= RS_MAXTOP, -3, "n_seg*k too large");
  int64_t grid = n_query < 148 * 64 ? n_query : 148 * 64;
  size_t smem = (size_t)dim * sizeof(double);
  KNN_CHECK_ARG(smem <= 32 * 1024, -3, "dim %d too large for the rescoring kernel", dim);
  knn_rescore_kernel<<<(unsigned)grid, RS_THREADS, smem, stream>>>(q, qn, n_query, p, pn, n_pool, dim, k, pl.n_seg,
                                                                   pl.cap, log_val, log_idx, log_cnt, seg_top,
                                                                   index_offset, out_dist, out_dist64, out_idx,
                                                                   flag_list, flag_count, stats, mask_lo, mask_hi,
                                                                   q_err, p_err);
  KNN_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- merge_topk (C1)
// One warp per query row; n_shards*k <= 512 entries staged in shared memory and
// ranked by (dist, idx).  Exact ties resolve to the lower global index, so the
// result is independent of how the pool was sharded.
constexpr int MG_WARPS = 4;
constexpr int MG_MAX = 512;

template <typename DT>
__global__ void __launch_bounds__(MG_WARPS * 32) merge_topk_kernel(const DT* __restrict__ gd,
                                                                   const int64_t* __restrict__ gi, int n_shards,
                                                                   int64_t n_query, int k, float* __restrict__ out_dist,
                                                                   DT* __restrict__ out_dist_full,
                                                                   int64_t* __restrict__ out_idx) {
  __shared__ DT sd[MG_WARPS][MG_MAX];
  __shared__ int64_t si[MG_WARPS][MG_MAX];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = n_shards * k;
  for (int64_t row = (int64_t)blockIdx.x * MG_WARPS + warp; row < n_query; row += (int64_t)gridDim.x * MG_WARPS) {
    __syncwarp();
    for (int e = lane; e < n; e += 32) {
      const int s = e / k, j = e % k;
      sd[warp][e] = gd[((int64_t)s * n_query + row) * k + j];
      si[warp][e] = gi[((int64_t)s * n_query + row) * k + j];
    }
    __syncwarp();
    for (int e = lane; e < n; e += 32) {
      const DT d = sd[warp][e];
      const int64_t i = si[warp][e];
      if (i < 0) continue;  // padding of a shard smaller than k
      int rank = 0;
      for (int j = 0; j < n; ++j) {
        const DT dj = sd[warp][j];
        const int64_t ij = si[warp][j];
        rank += (ij >= 0) && ((dj < d) || (dj == d && ij < i));
      }
      if (rank < k) {
        out_dist[row * k + rank] = (float)d;
        if (out_dist_full) out_dist_full[row * k + rank] = d;
        out_idx[row * k + rank] = i;
      }
    }
  }
}

int launch_merge_topk(const float* gd, const int64_t* gi, int n_shards, int64_t n_query, int k, float* out_dist,
                      int64_t* out_idx, cudaStream_t stream) {
  if (n_query == 0) return 0;
  KNN_CHECK_ARG(n_shards >= 1 && n_shards * k <= MG_MAX, -3, "merge_topk: n_shards*k=%d exceeds %d", n_shards * k,
                MG_MAX);
  int64_t grid = ceil_div64(n_query, MG_WARPS);
  if (grid > 148 * 16) grid = 148 * 16;
  merge_topk_kernel<float><<<(unsigned)grid, MG_WARPS * 32, 0, stream>>>(gd, gi, n_shards, n_query, k, out_dist,
                                                                          nullptr, out_idx);
  KNN_LAUNCH_CHECK();
  return 0;
}

// The same merge on the fp64 distances the re-score ranked by: a pool searched in N shards then
// returns bit for bit what one search of the whole pool returns (the fp32 roundings of two
// different fp64 distances can tie; ranking the shards' lists on the rounded values could then
// order, or cut, differently from the single search).
int launch_merge_topk64(const double* gd, const int64_t* gi, int n_shards, int64_t n_query, int k, float* out_dist,
                        double* out_dist64, int64_t* out_idx, cudaStream_t stream) {
  if (n_query == 0) return 0;
  KNN_CHECK_ARG(n_shards >= 1 && n_shards * k <= MG_MAX, -3, "merge_topk: n_shards*k=%d exceeds %d", n_shards * k,
                MG_MAX);
  int64_t grid = ceil_div64(n_query, MG_WARPS);
  if (grid > 148 * 16) grid = 148 * 16;
  merge_topk_kernel<double><<<(unsigned)grid, MG_WARPS * 32, 0, stream>>>(gd, gi, n_shards, n_query, k, out_dist,
                                                                           out_dist64, out_idx);
  KNN_LAUNCH_CHECK();
  return 0;
}

}  // namespace knnsvc
