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
// Survivors are streamed (no cap): each warp re-scores its share and keeps a
// sorted exact top-k; the warps' lists are merged at the end.  Rows whose log
// overflowed in some segment (more than `cap` candidates inside the window:
// massive ties) are appended to the flag list for the exact brute-force kernel.
#include <limits.h>

#include "common.cuh"
#include "kernels.cuh"

namespace knnsvc {

constexpr int RS_THREADS = 128;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_MAXTOP = 16 * kMaxK;  // n_seg <= 16

__device__ __forceinline__ bool rs_less(double da, int ia, double db, int ib) {
  return da < db || (da == db && ia < ib);
}

__global__ void __launch_bounds__(RS_THREADS) knn_rescore_kernel(
    const float* __restrict__ q, const float* __restrict__ qn, int64_t n_query, const float* __restrict__ p,
    const float* __restrict__ pn, int64_t n_pool, int dim, int k, int n_seg, int cap,
    const float* __restrict__ log_val, const int* __restrict__ log_idx, const int* __restrict__ log_cnt,
    const float* __restrict__ seg_top, int64_t index_offset, float* __restrict__ out_dist,
    int64_t* __restrict__ out_idx, int64_t* __restrict__ flag_list, int* __restrict__ flag_count,
    int* __restrict__ stats, const int64_t* __restrict__ mask_lo, const int64_t* __restrict__ mask_hi,
    const float* __restrict__ q_err, const float* __restrict__ p_err) {
  extern __shared__ __align__(16) float s_q[];  // [dim]
  __shared__ float s_top[RS_MAXTOP];
  __shared__ double s_d[RS_WARPS][kMaxK];   // per-warp exact top-k, ascending by (dist, idx)
  __shared__ int s_i[RS_WARPS][kMaxK];
  __shared__ int s_nsurv[RS_WARPS];
  __shared__ int s_overflow, s_logged;
  __shared__ float s_thr;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int64_t row = blockIdx.x; row < n_query; row += gridDim.x) {
    __syncthreads();
    if (tid == 0) {
      s_overflow = 0;
      s_logged = 0;
      s_thr = -INFINITY;
    }
    if (tid < RS_WARPS) s_nsurv[tid] = 0;
    const int n_top = n_seg * k;
    for (int e = tid; e < n_top; e += RS_THREADS) s_top[e] = seg_top[row * n_top + e];
    for (int c = tid; c < dim; c += RS_THREADS) s_q[c] = __ldg(q + row * dim + c);
    for (int e = tid; e < RS_WARPS * kMaxK; e += RS_THREADS) {
      s_d[e / kMaxK][e % kMaxK] = INFINITY;
      s_i[e / kMaxK][e % kMaxK] = INT_MAX;
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
      if (rank == k - 1) s_thr = v - 2.0f * filter_eps(q_err, p_err);
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
    const double qnorm = (double)qn[row];
    // masked column range: distance defined as 1 (ddsp_prematch_dataset.py:1623-1624)
    const int64_t m_lo = mask_lo ? mask_lo[row] : 0, m_hi = mask_lo ? mask_hi[row] : 0;
    int my_surv = 0;
    double kth_d = INFINITY;   // warp-uniform copy of this warp's current k-th best
    int kth_i = INT_MAX;
    for (int s = 0; s < n_seg; ++s) {
      const int64_t slot = row * n_seg + s;
      const int c = log_cnt[slot];
      for (int base = warp * 32; base < c; base += RS_THREADS) {
        const int e = base + lane;
        int cand = -1;
        if (e < c && log_val[slot * cap + e] >= thr) cand = log_idx[slot * cap + e];
        unsigned mask = __ballot_sync(0xffffffffu, cand >= 0);
        while (mask) {
          const int b = __ffs(mask) - 1;
          mask &= mask - 1;
          const int pr = __shfl_sync(0xffffffffu, cand, b);
          const float* prow = p + (int64_t)pr * dim;
          double acc = 0.0;
          const bool masked = pr >= m_lo && pr < m_hi;   // warp-uniform
          if (masked) {
          } else if ((dim & 3) == 0) {
            const float4* p4 = reinterpret_cast<const float4*>(prow);
            const float4* q4 = reinterpret_cast<const float4*>(s_q);
            for (int cc = lane; cc < dim / 4; cc += 32) {
              const float4 a = __ldg(p4 + cc);
              const float4 bq = q4[cc];
              acc += (double)a.x * bq.x + (double)a.y * bq.y + (double)a.z * bq.z + (double)a.w * bq.w;
            }
          } else {
            for (int cc = lane; cc < dim; cc += 32) acc += (double)__ldg(prow + cc) * (double)s_q[cc];
          }
          acc = warp_sum(acc);
          const double d = masked ? 1.0 : 1.0 - acc / (qnorm * (double)pn[pr]);
          ++my_surv;
          if (rs_less(d, pr, kth_d, kth_i)) {
            if (lane == 0) {
              int j = k - 1;
              while (j > 0 && rs_less(d, pr, s_d[warp][j - 1], s_i[warp][j - 1])) {
                s_d[warp][j] = s_d[warp][j - 1];
                s_i[warp][j] = s_i[warp][j - 1];
                --j;
              }
              s_d[warp][j] = d;
              s_i[warp][j] = pr;
            }
            __syncwarp();
            kth_d = s_d[warp][k - 1];
            kth_i = s_i[warp][k - 1];
          }
        }
      }
    }
    if (lane == 0) s_nsurv[warp] = my_surv;
    __syncthreads();
    // merge the RS_WARPS sorted lists: rank of each entry among all RS_WARPS*k by (dist, idx)
    for (int e = tid; e < RS_WARPS * k; e += RS_THREADS) {
      const int w = e / k, j = e % k;
      const double d = s_d[w][j];
      const int i = s_i[w][j];
      if (i == INT_MAX) continue;
      int rank = 0;
      for (int w2 = 0; w2 < RS_WARPS; ++w2)
        for (int j2 = 0; j2 < k; ++j2) rank += rs_less(s_d[w2][j2], s_i[w2][j2], d, i);
      if (rank < k) {
        out_dist[row * k + rank] = (float)d;
        out_idx[row * k + rank] = (int64_t)i + index_offset;
      }
    }
    if (tid == 0 && stats) {
      int ns = 0;
      for (int w = 0; w < RS_WARPS; ++w) ns += s_nsurv[w];
      atomicAdd(stats + 1, s_logged);
      atomicAdd(stats + 2, ns);
    }
  }
}

int launch_knn_rescore(const float* q, const float* qn, int64_t n_query, const float* p, const float* pn,
                       int64_t n_pool, int dim, int k, const FilterPlan& pl, const float* log_val,
                       const int* log_idx, const int* log_cnt, const float* seg_top, int64_t index_offset,
                       float* out_dist, int64_t* out_idx, int64_t* flag_list, int* flag_count, int* stats,
                       const int64_t* mask_lo, const int64_t* mask_hi, const float* q_err, const float* p_err,
                       cudaStream_t stream) {
  if (n_query == 0) return 0;
  KNN_CHECK_ARG(pl.n_seg * k <= RS_MAXTOP, -3, "n_seg*k too large");
  int64_t grid = n_query < 148 * 64 ? n_query : 148 * 64;
  size_t smem = (size_t)dim * sizeof(float);
  KNN_CHECK_ARG(smem <= 32 * 1024, -3, "dim %d too large for the rescoring kernel", dim);
  knn_rescore_kernel<<<(unsigned)grid, RS_THREADS, smem, stream>>>(q, qn, n_query, p, pn, n_pool, dim, k, pl.n_seg,
                                                                   pl.cap, log_val, log_idx, log_cnt, seg_top,
                                                                   index_offset, out_dist, out_idx, flag_list,
                                                                   flag_count, stats, mask_lo, mask_hi, q_err, p_err);
  KNN_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- merge_topk (C1)
// One warp per query row; n_shards*k <= 512 entries staged in shared memory and
// ranked by (dist, idx).  Exact ties resolve to the lower global index, so the
// result is independent of how the pool was sharded.
constexpr int MG_WARPS = 4;
constexpr int MG_MAX = 512;

__global__ void __launch_bounds__(MG_WARPS * 32) merge_topk_kernel(const float* __restrict__ gd,
                                                                   const int64_t* __restrict__ gi, int n_shards,
                                                                   int64_t n_query, int k, float* __restrict__ out_dist,
                                                                   int64_t* __restrict__ out_idx) {
  __shared__ float sd[MG_WARPS][MG_MAX];
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
      const float d = sd[warp][e];
      const int64_t i = si[warp][e];
      if (i < 0) continue;  // padding of a shard smaller than k
      int rank = 0;
      for (int j = 0; j < n; ++j) {
        const float dj = sd[warp][j];
        const int64_t ij = si[warp][j];
        rank += (ij >= 0) && ((dj < d) || (dj == d && ij < i));
      }
      if (rank < k) {
        out_dist[row * k + rank] = d;
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
  merge_topk_kernel<<<(unsigned)grid, MG_WARPS * 32, 0, stream>>>(gd, gi, n_shards, n_query, k, out_dist, out_idx);
  KNN_LAUNCH_CHECK();
  return 0;
}

}  // namespace knnsvc
