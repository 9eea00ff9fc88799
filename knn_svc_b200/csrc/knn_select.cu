// Exact re-scoring of the tensor-core filter's candidate log, and the top-k
// merge used after the NCCL all-gather of per-shard results (C1).
//
// Three kernels turn a row's log into its exact top-k:
//   knn_refine_prep  (one warp per row)  tau* = k-th largest approximate similarity over the row's
//      per-segment top-k lists -> the row's first threshold tau* - 2*eps (every true top-k member has
//      s~ >= tau* - 2*eps); rows whose log overflowed go to the flag list (exact brute force decides);
//      and, because every (row, segment) log is SORTED by pool column (the filter walks the pool
//      in order), the offsets at which each refine block of pool rows starts in it.
//   knn_refine       (block-major stream) every candidate above the first threshold is scored again
//      with an fp32 dot product of the fp32 rows (query row in registers, pool row read once,
//      coalesced).  Work is ordered by POOL BLOCK (64 MB of rows) and inside a block by query row,
//      so that all warps gather from the same L2-resident block: dense (WavLM-like) data has
//      ~10^3 candidates per row and the gather, not arithmetic, is what the stage costs.
//      |s1 - s| <= eps1 ~ 1e-6 (rigorous: fma chains of dim/128 + a 7-level tree, see refine_eps),
//      600x tighter than the fp16 window.
//   knn_rescore      (one CTA per row) tau1 = k-th largest s1; the few candidates with
//      s1 >= tau1 - 2*eps1 (same superset argument) are scored from the fp32 rows with fp64
//      accumulation,  dist = 1 - q.p/(|q||p|)  (lib_ongaku_test.py:162-165), ranked by
//      (dist, index) and the first k written ascending — the order dists.topk(k, largest=False)
//      returns (ddsp_prematch_dataset.py:1203).
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

// fp64 dot products of ONE or TWO pool rows with a query row (products of two fp32 values are exact
// in fp64; the element-to-lane assignment, the two fma chains per lane and the reduction tree are
// fixed, and rows.cu's exact kernel follows the same order, so every kernel that scores a pair gets
// the same bits).  Two rows per pass share the query loads and give two independent DFMA chains.
template <bool VEC>
__device__ __forceinline__ void rs_dot2(const float* __restrict__ pa, const float* __restrict__ pb,
                                        const float* __restrict__ qrow, int dim, int lane, double& ra, double& rb) {
  double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
  if (VEC) {
    const float4* a4 = reinterpret_cast<const float4*>(pa);
    const float4* b4 = reinterpret_cast<const float4*>(pb);
    const float4* q4 = reinterpret_cast<const float4*>(qrow);
    const int nv = dim / 4;
#pragma unroll 4
    for (int c = lane; c < nv; c += 32) {
      const float4 va = __ldg(a4 + c);
      const float4 vb = __ldg(b4 + c);
      const float4 qv = __ldg(q4 + c);
      a0 = fma((double)va.x, (double)qv.x, a0);
      a1 = fma((double)va.y, (double)qv.y, a1);
      a0 = fma((double)va.z, (double)qv.z, a0);
      a1 = fma((double)va.w, (double)qv.w, a1);
      b0 = fma((double)vb.x, (double)qv.x, b0);
      b1 = fma((double)vb.y, (double)qv.y, b1);
      b0 = fma((double)vb.z, (double)qv.z, b0);
      b1 = fma((double)vb.w, (double)qv.w, b1);
    }
  } else {
    for (int c = lane; c < dim; c += 32) {
      const double qv = (double)__ldg(qrow + c);
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

// ---------------------------------------------------------------- refine: plan
// Refine blocks: rf_rows pool rows each (64 MB of fp32 rows at 1024 dims: half of L2), at most
// kRefineMaxBlocks of them (the per-(row, segment) offset table has rf_nblk + 1 entries).
constexpr int kRefineMaxBlocks = 1024;
// Route of the decision stage, chosen ON THE DEVICE, PER QUERY ROW, from what the prep kernel counts:
//   refine  more than kRefineMinAbove candidates above the row's first (fp16-window) threshold — pools of
//           near-duplicates: they are first re-scored in fp32, block-major (knn_refine), and only the few
//           inside the refined window are scored in fp64;
//   direct  otherwise: the candidates above the first threshold go straight to the fp64 scoring — by one
//           warp when there are at most 32 of them, else by one CTA per row (adjacent rows run side by side
//           on an SM and share their candidates through L1).  Measured (profiles/r2_dense_search_*.jsonl):
//           with ~30-150 candidates per row the extra pass over the log costs more than the fp64
//           arithmetic it saves (100k x 30k, k=32: 5.5 ms direct, 9.2 ms refined), with ~800-1100 it wins
//           (20k x 1M near-duplicate pool, k=4: 14.3 -> 9.5 ms).
constexpr int kRefineMinAbove = 400;
 // per-row route codes (row_mode)
constexpr int kRouteDirectCta = 0;       // > 32 candidates above the first threshold, scored in fp64 by one CTA
constexpr int kRouteRefine = 1;          // fp32 re-score first
constexpr int kRouteDirectWarp = 2;      // <= 32 candidates above the first threshold, scored in fp64 by one warp
constexpr int kStatAbove = 8;            // stats slot (relative to the stats base): number of rows on the refine route
void plan_refine(int64_t n_query, int64_t n_pool, int dim, int* rf_rows, int* rf_nblk) {
  int64_t rows = (int64_t)(64 << 20) / ((int64_t)dim * 4);
  rows = rows < 256 ? 256 : rows / 256 * 256;
  int64_t nb = ceil_div64(n_pool, rows);
  // few query rows: cut the pool finer so that there is a warp unit (block x 32-row group) for every
  // resident warp of the GPU (148 SMs x 16 warps, twice over)
  const int64_t groups = ceil_div64(n_query < 1 ? 1 : n_query, 32);
  const int64_t want = ceil_div64(2 * 148 * 16, groups);
  if (nb < want) nb = want;
  if (nb > kRefineMaxBlocks) nb = kRefineMaxBlocks;
  rows = ceil_div64(ceil_div64(n_pool, nb), 256) * 256;
  nb = ceil_div64(n_pool, rows);
  *rf_rows = (int)rows;
  *rf_nblk = (int)(nb < 1 ? 1 : nb);
}

// fp32 refine error bound in cosine units: every product enters a chain of `chain` fmas, then a
// 2-level in-thread tree and a 5-level shuffle tree; |fl(sum) - sum| <= depth * 2^-24 * sum|q_i p_i|
// <= depth * 2^-24 * |q||p|.  + one rounding of the quotient to fp32, + 10% slack.
__host__ __device__ inline float refine_eps(int dim, bool vec) {
  const int chain = vec ? (dim + 127) / 128 : (dim + 31) / 32;
  return (float)(chain + 2 + 5 + 1) * 5.9604645e-8f * 1.1f;
}

constexpr int RP_WARPS = 4;
constexpr int RP_MAXTOP = 16 * kMaxK;

// One warp per query row.
__global__ void __launch_bounds__(RP_WARPS * 32) knn_refine_prep_kernel(
    int64_t n_query, int k, int n_seg, int cap, int rf_rows, int rf_nblk, const int* __restrict__ log_idx,
    const int* __restrict__ log_cnt, const float* __restrict__ seg_top, float* __restrict__ row_thr,
    int* __restrict__ blk_off, int64_t* __restrict__ flag_list, int* __restrict__ flag_count, int* __restrict__ stats,
    const float* __restrict__ log_val, const float* __restrict__ q_err, const float* __restrict__ p_err, int dim_pad,
    int min_above, int* __restrict__ row_mode) {
  __shared__ float s_top[RP_WARPS][RP_MAXTOP];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_top = n_seg * k;
  const float window = 2.0f * filter_eps(q_err, p_err, dim_pad);
  int logged = 0, n_refine = 0;
  for (int64_t row = (int64_t)blockIdx.x * RP_WARPS + warp; row < n_query; row += (int64_t)gridDim.x * RP_WARPS) {
    __syncwarp();
    for (int e = lane; e < n_top; e += 32) s_top[warp][e] = seg_top[row * n_top + e];
    bool overflow = false;
    for (int s = lane; s < n_seg; s += 32) {
      const int c = log_cnt[row * n_seg + s];
      overflow |= c > cap;
      logged += c > cap ? cap : c;
    }
    overflow = __any_sync(0xffffffffu, overflow);
    __syncwarp();
    // k-th largest of the union of the segment lists (rank counting, ties by position)
    float thr = INFINITY;
    for (int e = lane; e < n_top; e += 32) {
      const float v = s_top[warp][e];
      int rank = 0;
      for (int j = 0; j < n_top; ++j) {
        const float o = s_top[warp][j];
        rank += (o > v) || (o == v && j < e);
      }
      if (rank == k - 1) thr = v - window;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) thr = fminf(thr, __shfl_xor_sync(0xffffffffu, thr, o));
    if (overflow) {
      // more than `cap` candidates inside the error window of one segment: exact brute force decides
      if (lane == 0) {
        flag_list[atomicAdd(flag_count, 1)] = row;
        if (stats) atomicAdd(stats + 0, 1);
        row_thr[row] = INFINITY;
        row_mode[row] = kRouteDirectCta;
      }
      continue;
    }
    // candidates above the first threshold -> the row's route
    int above = 0;
    for (int s = 0; s < n_seg; ++s) {
      const int64_t slot = row * n_seg + s;
      const int c = log_cnt[slot];
      const float* lvp = log_val + slot * cap;
      for (int e = lane; e < c; e += 32) above += lvp[e] >= thr ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) above += __shfl_xor_sync(0xffffffffu, above, o);
    const bool refine = above > min_above;
    if (lane == 0) {
      row_thr[row] = thr;
      row_mode[row] = refine ? kRouteRefine : (above <= 32 ? kRouteDirectWarp : kRouteDirectCta);
    }
    if (!refine) continue;              // direct route: no block offsets needed
    ++n_refine;
    // block offsets of every (row, segment) log: off[b] = first entry whose column lies in block >= b
    for (int s = 0; s < n_seg; ++s) {
      const int64_t slot = row * n_seg + s;
      const int c = log_cnt[slot];
      const int* li = log_idx + slot * cap;
      // table layout [block][slot]: the refine kernel reads one block's offsets of 32 consecutive rows at once
      const int64_t n_slots = n_query * n_seg;
      int* off = blk_off + slot;
      for (int e0 = 0; e0 < c; e0 += 32) {
        const int e = e0 + lane;
        const int b = e < c ? li[e] / rf_rows : 0;
        int pb = __shfl_up_sync(0xffffffffu, b, 1);
        if (lane == 0) pb = e0 > 0 ? li[e0 - 1] / rf_rows : -1;
        if (e < c)
          for (int j = pb + 1; j <= b; ++j) off[(int64_t)j * n_slots] = e;
      }
      const int last_b = c > 0 ? li[c - 1] / rf_rows : -1;
      for (int j = last_b + 1 + lane; j <= rf_nblk; j += 32) off[(int64_t)j * n_slots] = c;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) logged += __shfl_xor_sync(0xffffffffu, logged, o);
  if (lane == 0 && stats && logged) atomicAdd(stats + 1, logged);
  if (lane == 0 && stats && n_refine) atomicAdd(stats + kStatAbove, n_refine);
}

// fp32 dot products of TWO pool rows with the query row held in registers (NV float4 per lane,
// element 4*(lane + 32*j) .. +3): independent fma chains per component, a fixed reduction order.
template <int NV>
__device__ __forceinline__ void refine_dot2(const float4 (&qv)[NV], const float* __restrict__ pa,
                                            const float* __restrict__ pb, int lane, float& ra, float& rb) {
  const float4* a4 = reinterpret_cast<const float4*>(pa);
  const float4* b4 = reinterpret_cast<const float4*>(pb);
  float4 va[NV], vb[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    va[j] = __ldg(a4 + lane + 32 * j);
    vb[j] = __ldg(b4 + lane + 32 * j);
  }
  float ax = 0.f, ay = 0.f, az = 0.f, aw = 0.f, bx = 0.f, by = 0.f, bz = 0.f, bw = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    ax = fmaf(va[j].x, qv[j].x, ax);
    ay = fmaf(va[j].y, qv[j].y, ay);
    az = fmaf(va[j].z, qv[j].z, az);
    aw = fmaf(va[j].w, qv[j].w, aw);
    bx = fmaf(vb[j].x, qv[j].x, bx);
    by = fmaf(vb[j].y, qv[j].y, by);
    bz = fmaf(vb[j].z, qv[j].z, bz);
    bw = fmaf(vb[j].w, qv[j].w, bw);
  }
  ra = (ax + ay) + (az + aw);
  rb = (bx + by) + (bz + bw);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ra += __shfl_xor_sync(0xffffffffu, ra, o);
    rb += __shfl_xor_sync(0xffffffffu, rb, o);
  }
}

// Block-major refine.  A warp unit is (refine block b, group of 32 consecutive query rows): lane l
// looks up the entries of row 32*g + l that fall into block b (two offset loads, coalesced over the
// group), and the warp then scores them row by row.  Units are dealt to warps in block-major order,
// so at any time the whole grid gathers from one or two blocks of the pool.
//   NV > 0: dim == NV * 128, rows 16-byte aligned: query row in registers.   NV == 0: any dim.
template <int NV>
__global__ void __launch_bounds__(256, 3) knn_refine_kernel(
    const float* __restrict__ q, const double* __restrict__ qn, int64_t n_query, const float* __restrict__ p,
    const double* __restrict__ pn, int dim, int n_seg, int cap, int rf_nblk, const float* __restrict__ log_val,
    const int* __restrict__ log_idx, const int* __restrict__ row_mode, const int* __restrict__ blk_off,
    const float* __restrict__ row_thr, float* __restrict__ ref_val, int* __restrict__ stats,
    const int64_t* __restrict__ mask_lo, const int64_t* __restrict__ mask_hi) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t n_groups = ceil_div64(n_query, 32);
  if (stats[kStatAbove] == 0) return;      // no row of this search is on the refine route
  const int64_t n_units = n_groups * rf_nblk;
  int scored = 0;
  for (int64_t u = warp; u < n_units; u += n_warps) {
    const int b = (int)(u / n_groups);
    const int64_t g = u - (int64_t)b * n_groups;
    const int64_t my_row = g * 32 + lane;
    const float my_thr = my_row < n_query ? row_thr[my_row] : INFINITY;
    const bool my_mode = my_row < n_query && row_mode[my_row] == kRouteRefine;
    if (!__any_sync(0xffffffffu, my_mode)) continue;
    for (int s = 0; s < n_seg; ++s) {
      int o0 = 0, o1 = 0;
      if (my_row < n_query && my_mode) {
        const int64_t n_slots = n_query * n_seg;
        const int* off = blk_off + (int64_t)b * n_slots + (my_row * n_seg + s);
        o0 = off[0];
        o1 = off[n_slots];
      }
      // lane-parallel pre-check for rows with a handful of entries (sparse data logs ~1 entry per row
      // and block, few of which pass): their entries below the first threshold are settled here and
      // only rows with a passing entry enter the cooperative loop below
      bool any_pass = o1 - o0 > 4;
      if (o1 > o0 && !any_pass) {
        const float* lv0 = log_val + (my_row * n_seg + s) * cap;
        float* rv0 = ref_val + (my_row * n_seg + s) * cap;
        for (int e = o0; e < o1; ++e) {
          if (lv0[e] >= my_thr) any_pass = true;
          else rv0[e] = -INFINITY;
        }
      }
      unsigned rows_mask = __ballot_sync(0xffffffffu, any_pass);
      while (rows_mask) {
        const int src = __ffs(rows_mask) - 1;
        rows_mask &= rows_mask - 1;
        const int64_t row = g * 32 + src;
        const int e_lo = __shfl_sync(0xffffffffu, o0, src), e_hi = __shfl_sync(0xffffffffu, o1, src);
        const float thr = __shfl_sync(0xffffffffu, my_thr, src);
        const int64_t slot = row * n_seg + s;
        const float* lv = log_val + slot * cap;
        const int* li = log_idx + slot * cap;
        float* rv = ref_val + slot * cap;
        const double qnorm = qn[row];
        const int64_t m_lo = mask_lo ? mask_lo[row] : 0, m_hi = mask_lo ? mask_hi[row] : 0;
        const float* qrow = q + row * dim;
        float4 qv[NV > 0 ? NV : 1];
        bool q_loaded = false;      // warp-uniform: the row is only fetched once one of its entries passes
        for (int e0 = e_lo; e0 < e_hi; e0 += 32) {
          const int e = e0 + lane;
          float sv = -INFINITY;
          int col = 0;
          if (e < e_hi) {
            sv = lv[e];
            col = li[e];
          }
          const bool pass = e < e_hi && sv >= thr;
          if (e < e_hi && !pass) rv[e] = -INFINITY;      // below the first threshold: cannot be a neighbour
          unsigned m = __ballot_sync(0xffffffffu, pass);
          if (lane == 0) scored += __popc(m);
          if (NV > 0 && m && !q_loaded) {
#pragma unroll
            for (int j = 0; j < (NV > 0 ? NV : 1); ++j)
              qv[j] = __ldg(reinterpret_cast<const float4*>(qrow) + lane + 32 * j);
            q_loaded = true;
          }
          while (m) {
            const int la = __ffs(m) - 1;
            m &= m - 1;
            const int lb = m ? __ffs(m) - 1 : la;
            m &= m - 1;                                  // (no-op when m was already 0)
            const int ca = __shfl_sync(0xffffffffu, col, la), cb = __shfl_sync(0xffffffffu, col, lb);
            // the two pool norms are fetched NOW (lanes 0 and 1), together with the rows, not after the sums
            const double pnorm = lane < 2 ? pn[lane == 0 ? ca : cb] : 1.0;
            float da, db;
            if (NV > 0) {
              refine_dot2<(NV > 0 ? NV : 1)>(qv, p + (int64_t)ca * dim, p + (int64_t)cb * dim, lane, da, db);
            } else {
              const float* pa = p + (int64_t)ca * dim;
              const float* pb = p + (int64_t)cb * dim;
              da = 0.f;
              db = 0.f;
              for (int c = lane; c < dim; c += 32) {
                const float qc = __ldg(qrow + c);
                da = fmaf(__ldg(pa + c), qc, da);
                db = fmaf(__ldg(pb + c), qc, db);
              }
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) {
                da += __shfl_xor_sync(0xffffffffu, da, o);
                db += __shfl_xor_sync(0xffffffffu, db, o);
              }
            }
            // masked column range: distance defined as 1 <=> similarity 0 (ddsp_prematch_dataset.py:1623-1624)
            if (lane == 0) rv[e0 + la] = (ca >= m_lo && ca < m_hi) ? 0.f : (float)((double)da / (qnorm * pnorm));
            if (lane == 1 && lb != la) rv[e0 + lb] = (cb >= m_lo && cb < m_hi) ? 0.f : (float)((double)db / (qnorm * pnorm));
          }
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) scored += __shfl_xor_sync(0xffffffffu, scored, o);
  if (lane == 0 && stats && scored) atomicAdd(stats + 2, scored);
}


// Rows with at most 32 candidates above the first threshold (all of sparse data at small k) are decided
// by ONE WARP each — candidates in registers, one per lane.  Refine-route rows: the k-th largest refined
// similarity by rank counting over shuffles, then fp64 scores of the survivors; direct-route rows: fp64
// scores of all of them (same rs_dot2); final (dist, index) rank.
// Rows it decides are marked (row_thr = +inf) and skipped by the CTA-per-row kernel that follows.
__global__ void __launch_bounds__(256) knn_rescore_small_kernel(
    const float* __restrict__ q, const double* __restrict__ qn, int64_t n_query, const float* __restrict__ p,
    const double* __restrict__ pn, int dim, int k, int n_seg, int cap, const float* __restrict__ ref_val,
    const float* __restrict__ log_val, const int* __restrict__ row_mode, const int* __restrict__ log_idx,
    const int* __restrict__ log_cnt, float* __restrict__ row_thr, float refine_window,
    int64_t index_offset, float* __restrict__ out_dist, double* __restrict__ out_dist64, int64_t* __restrict__ out_idx,
    int* __restrict__ stats, const int64_t* __restrict__ mask_lo, const int64_t* __restrict__ mask_hi) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const bool vec = (dim & 3) == 0 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(q) & 15) == 0);
  int n_scored = 0, n_direct = 0;
  for (int64_t row = warp; row < n_query; row += n_warps) {
    const float thr0 = row_thr[row];
    if (!(thr0 < INFINITY)) continue;                  // log overflow: the exact kernel decides (warp-uniform)
    const int route = row_mode[row];                   // warp-uniform
    if (route == kRouteDirectCta) continue;            // known to have more than 32 candidates
    const bool direct = route != kRouteRefine;
    // direct route: the approximate similarities against the first threshold; refine route: the refined
    // values (-inf below the first threshold)
    const float* __restrict__ vals = direct ? log_val : ref_val;
    const float cut = direct ? thr0 : -INFINITY;
    // gather the candidates above the first threshold, one per lane
    float myv = -INFINITY;
    int mycol = -1, n_valid = 0;
    bool too_many = false;
    for (int s = 0; s < n_seg && !too_many; ++s) {
      const int64_t slot = row * n_seg + s;
      int c = log_cnt[slot];
      c = c > cap ? cap : c;
      for (int e0 = 0; e0 < c; e0 += 32) {
        const int e = e0 + lane;
        const float v = e < c ? __ldg(vals + slot * cap + e) : -INFINITY;
        const unsigned m = __ballot_sync(0xffffffffu, direct ? (e < c && v >= cut) : v > -INFINITY);
        if (!m) continue;
        const int add = __popc(m);
        if (n_valid + add > 32) {
          too_many = true;
          break;
        }
        const int col = e < c ? log_idx[slot * cap + e] : -1;
        // lane t in [n_valid, n_valid + add) takes the (t - n_valid)-th set lane of m
        const int want = lane - n_valid;
        const int srcl = (want >= 0 && want < add) ? (int)__fns(m, 0, want + 1) : 0;
        const float tv = __shfl_sync(0xffffffffu, v, srcl);
        const int tc = __shfl_sync(0xffffffffu, col, srcl);
        if (want >= 0 && want < add) {
          myv = tv;
          mycol = tc;
        }
        n_valid += add;
      }
    }
    if (too_many) continue;                             // the CTA-per-row kernel takes this row
    // tau1 = k-th largest refined similarity (rank counting, ties by lane)
    int rnk = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float vj = __shfl_sync(0xffffffffu, myv, j);
      rnk += (j < n_valid) && ((vj > myv) || (vj == myv && j < lane));
    }
    const unsigned kth = __ballot_sync(0xffffffffu, lane < n_valid && rnk == k - 1);
    const float tau1 = (kth && !direct) ? __shfl_sync(0xffffffffu, myv, __ffs(kth) - 1) : -INFINITY;
    const bool surv = lane < n_valid && (direct || myv >= tau1 - refine_window);
    unsigned sm = __ballot_sync(0xffffffffu, surv);
    n_scored += __popc(sm);
    if (direct) n_direct += __popc(sm);
    // fp64 scores of the survivors, two per pass
    const float* qrow = q + row * dim;
    const double qnorm = qn[row];
    const int64_t m_lo = mask_lo ? mask_lo[row] : 0, m_hi = mask_lo ? mask_hi[row] : 0;
    double myd = INFINITY;
    while (sm) {
      const int la = __ffs(sm) - 1;
      sm &= sm - 1;
      const int lb = sm ? __ffs(sm) - 1 : la;
      sm &= sm - 1;
      const int ia = __shfl_sync(0xffffffffu, mycol, la), ib = __shfl_sync(0xffffffffu, mycol, lb);
      double da, db;
      if (vec) rs_dot2<true>(p + (int64_t)ia * dim, p + (int64_t)ib * dim, qrow, dim, lane, da, db);
      else rs_dot2<false>(p + (int64_t)ia * dim, p + (int64_t)ib * dim, qrow, dim, lane, da, db);
      // (rs_dot2 leaves the full sums in every lane)
      if (lane == la) myd = (ia >= m_lo && ia < m_hi) ? 1.0 : 1.0 - da / (qnorm * pn[ia]);
      if (lane == lb && lb != la) myd = (ib >= m_lo && ib < m_hi) ? 1.0 : 1.0 - db / (qnorm * pn[ib]);
    }
    // final order by (dist, index) among the survivors
    int fr = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const double dj = __shfl_sync(0xffffffffu, myd, j);
      const int cj = __shfl_sync(0xffffffffu, mycol, j);
      const bool sj = (j < n_valid) && dj < INFINITY;
      fr += sj && rs_less(dj, cj, myd, mycol);
    }
    if (surv && fr < k) {
      out_dist[row * k + fr] = (float)myd;
      if (out_dist64) out_dist64[row * k + fr] = myd;
      out_idx[row * k + fr] = (int64_t)mycol + index_offset;
    }
    __syncwarp();
    if (lane == 0) row_thr[row] = INFINITY;             // decided
  }
  // (both counts are warp-uniform: every lane holds its warp's total)
  if (lane == 0 && stats && n_scored) atomicAdd(stats + 6, n_scored);
  if (lane == 0 && stats && n_direct) atomicAdd(stats + 2, n_direct);
}

// order-preserving map float -> unsigned (larger float <=> larger key) and back
__device__ __forceinline__ unsigned float_key(float v) {
  const unsigned b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_float(unsigned key) {
  return __uint_as_float((key & 0x80000000u) ? (key & 0x7fffffffu) : ~key);
}

__global__ void __launch_bounds__(RS_THREADS) knn_rescore_kernel(
    const float* __restrict__ q, const double* __restrict__ qn, int64_t n_query, const float* __restrict__ p,
    const double* __restrict__ pn, int64_t n_pool, int dim, int k, int n_seg, int cap,
    const float* __restrict__ ref_val_in, const float* __restrict__ log_val, const int* __restrict__ log_idx,
    const int* __restrict__ log_cnt, const float* __restrict__ row_thr, const int* __restrict__ row_mode,
    float refine_window, int64_t index_offset, float* __restrict__ out_dist, double* __restrict__ out_dist64,
    int64_t* __restrict__ out_idx, int* __restrict__ stats, const int64_t* __restrict__ mask_lo,
    const int64_t* __restrict__ mask_hi) {
  __shared__ double s_dist[RS_LIST];             // exact distances of the survivors in s_cand
  __shared__ int s_cand[RS_LIST];
  __shared__ double s_best_d[kMaxK];             // scratch for the reduction to the best k
  __shared__ int s_best_i[kMaxK];
  __shared__ int s_hist[256];
  __shared__ int s_sel[2];
  __shared__ int s_n, s_nsurv;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool vec = (dim & 3) == 0 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(q) & 15) == 0);

  for (int64_t row = blockIdx.x; row < n_query; row += gridDim.x) {
    if (!(row_thr[row] < INFINITY)) continue;    // log overflow / decided by the warp kernel (block-uniform)
    // direct route: the candidates above the FIRST threshold are scored in fp64 right away (no refined values exist)
    const bool direct = row_mode[row] != kRouteRefine;      // block-uniform
    const float* __restrict__ ref_val = direct ? log_val : ref_val_in;
    const float* qrow = q + row * dim;
    __syncthreads();
    if (tid == 0) {
      s_nsurv = 0;
      s_n = 0;
    }
    // tau1 = k-th largest refined similarity of the row (with multiplicity): a 4-pass radix select
    // over the order-preserving integer image of the floats, 8 bits per pass, 256-bin histogram in
    // shared memory.  Entries below the first threshold hold -inf and are skipped.  Cost does not
    // depend on k; the entries (a few KB) stay in L1 between the passes.
    unsigned prefix = 0, prefix_mask = 0;
    int want = k;                  // rank still to find inside the current prefix bucket (1-based, from the top)
    bool have_tau = true;
    for (int pass = 0; pass < (direct ? 0 : 4); ++pass) {
      const int shift = 24 - 8 * pass;
      for (int i = tid; i < 256; i += RS_THREADS) s_hist[i] = 0;
      __syncthreads();
      for (int s = 0; s < n_seg; ++s) {
        const int64_t slot = row * n_seg + s;
        const int c = log_cnt[slot];
        const float* rv = ref_val + slot * cap;
        for (int e = tid; e < c; e += RS_THREADS) {
          const float v = __ldg(rv + e);
          if (!(v > -INFINITY)) continue;
          const unsigned key = float_key(v);
          if ((key & prefix_mask) == prefix) atomicAdd(&s_hist[(key >> shift) & 255u], 1);
        }
      }
      __syncthreads();
      // walk the bins from the top: the bin in which the running count reaches `want`
      if (warp == 0) {
        int carry = 0, found_bin = -1, found_before = 0;
        for (int base = 255; base >= 0 && found_bin < 0; base -= 32) {
          const int bin = base - lane;                       // lane 0 = highest bin of this group of 32
          const int cnt = s_hist[bin];
          int incl = cnt;                                    // inclusive prefix over lanes 0..lane
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
          }
          const bool hit = carry + incl >= want && carry + incl - cnt < want;
          const unsigned hm = __ballot_sync(0xffffffffu, hit);
          if (hm) {
            const int hl = __ffs(hm) - 1;
            found_bin = base - hl;
            found_before = carry + __shfl_sync(0xffffffffu, incl - cnt, hl);
          }
          carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) {
          s_sel[0] = found_bin;
          s_sel[1] = found_before;
        }
      }
      __syncthreads();
      const int bin = s_sel[0];
      if (bin < 0) {               // fewer than k candidates above the first threshold: everything survives
        have_tau = false;
        break;
      }
      want -= s_sel[1];
      prefix |= (unsigned)bin << shift;
      prefix_mask |= 255u << shift;
    }
    const float prev_v = have_tau ? key_float(prefix) : -INFINITY;
    const float thr = direct ? row_thr[row] : (prev_v > -INFINITY ? prev_v - refine_window : -INFINITY);   // block-uniform
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
        if (vec) rs_dot2<true>(p + (int64_t)ia * dim, p + (int64_t)ib * dim, qrow, dim, lane, da, db);
        else rs_dot2<false>(p + (int64_t)ia * dim, p + (int64_t)ib * dim, qrow, dim, lane, da, db);
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

    __syncthreads();   // s_n = 0 is visible
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
        if (e < c && __ldg(ref_val + slot * cap + e) >= thr) {
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
      atomicAdd(stats + 6, s_nsurv);
      if (direct) atomicAdd(stats + 2, s_nsurv);
    }
  }
}

int launch_knn_rescore(const float* q, const double* qn, int64_t n_query, const float* p, const double* pn,
                       int64_t n_pool, int dim, int k, const FilterPlan& pl, const float* log_val,
                       const int* log_idx, const int* log_cnt, const float* seg_top, float* ref_val, float* row_thr,
                       int* row_mode, int* blk_off, int64_t index_offset, float* out_dist, double* out_dist64, int64_t* out_idx,
                       int64_t* flag_list, int* flag_count, int* stats, const int64_t* mask_lo, const int64_t* mask_hi,
                       const float* q_err, const float* p_err, cudaStream_t stream) {
  if (n_query == 0) return 0;
  KNN_CHECK_ARG(pl.n_seg * k <= RP_MAXTOP, -3, "n_seg*k too large");
  const int dim_pad = (dim + 63) / 64 * 64;
  // 1. thresholds, overflow flags, block offsets
  {
    int64_t grid = ceil_div64(n_query, RP_WARPS);
    if (grid > 148 * 16) grid = 148 * 16;
    knn_refine_prep_kernel<<<(unsigned)grid, RP_WARPS * 32, 0, stream>>>(n_query, k, pl.n_seg, pl.cap, pl.rf_rows,
                                                                       pl.rf_nblk, log_idx, log_cnt, seg_top, row_thr,
                                                                       blk_off, flag_list, flag_count, stats, log_val,
                                                                       q_err, p_err, dim_pad,
                                                                       opt_refine_min() > 0 ? opt_refine_min() : kRefineMinAbove,
                                                                       row_mode);
    KNN_LAUNCH_CHECK();
  }
  // 2. fp32 refine of every candidate above the first threshold, block-major
  const bool vec = dim % 128 == 0 && dim <= 1024 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(q) & 15) == 0);
  {
    const int64_t units = ceil_div64(n_query, 32) * pl.rf_nblk;
    int64_t grid = ceil_div64(units, 8);
    const int64_t max_grid = (int64_t)148 * 8;        // units are dealt to warps in block-major order, a few waves
    if (grid > max_grid) grid = max_grid;
#define KNN_REFINE_GO(NV)                                                                                          \
  knn_refine_kernel<NV><<<(unsigned)grid, 256, 0, stream>>>(q, qn, n_query, p, pn, dim, pl.n_seg, pl.cap, pl.rf_nblk, \
                                                            log_val, log_idx, row_mode, blk_off, row_thr, ref_val,   \
                                                            stats, mask_lo, mask_hi)
    switch (vec ? dim / 128 : 0) {
      case 1: KNN_REFINE_GO(1); break;
      case 2: KNN_REFINE_GO(2); break;
      case 3: KNN_REFINE_GO(3); break;
      case 4: KNN_REFINE_GO(4); break;
      case 5: KNN_REFINE_GO(5); break;
      case 6: KNN_REFINE_GO(6); break;
      case 7: KNN_REFINE_GO(7); break;
      case 8: KNN_REFINE_GO(8); break;
      default: KNN_REFINE_GO(0); break;
    }
#undef KNN_REFINE_GO
    KNN_LAUNCH_CHECK();
  }
  // 3. fp64 decision among the candidates inside the refined window: rows with <= 32 candidates above the
  //    first threshold by one warp each, the others by one CTA each
  {
    int64_t g2 = ceil_div64(n_query, 8);
    if (g2 > 148 * 8) g2 = 148 * 8;
    knn_rescore_small_kernel<<<(unsigned)g2, 256, 0, stream>>>(q, qn, n_query, p, pn, dim, k, pl.n_seg, pl.cap, ref_val,
                                                             log_val, row_mode, log_idx, log_cnt, row_thr,
                                                             2.0f * refine_eps(dim, vec),
                                                             index_offset, out_dist, out_dist64, out_idx, stats, mask_lo,
                                                             mask_hi);
    KNN_LAUNCH_CHECK();
  }
  int64_t grid = n_query < 148 * 64 ? n_query : 148 * 64;
  knn_rescore_kernel<<<(unsigned)grid, RS_THREADS, 0, stream>>>(q, qn, n_query, p, pn, n_pool, dim, k, pl.n_seg,
                                                                   pl.cap, ref_val, log_val, log_idx, log_cnt, row_thr,
                                                                   row_mode, 2.0f * refine_eps(dim, vec), index_offset, out_dist,
                                                                   out_dist64, out_idx, stats, mask_lo, mask_hi);
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
