// K5, staged variants: greedy concatenation-cost re-selection with every row the
// recurrence touches resident in shared memory BEFORE the step that needs it.
// knn_with_concat_cost — lib_ongaku_test.py:270-369, K = 4.
//
// Two kernels, same bits: concat_cost_staged_kernel (one CTA per utterance: batches) directly below, and
// concat_cost_cluster_kernel (one cluster of 8 CTAs per utterance, the feature dimension split over the
// cluster: launches of up to 15 utterances) further down; the arithmetic of a step lives in the cs_* functions
// both call.  post.cu holds the general kernel for row shapes neither takes.
//
// The recurrence is serial in the frame index, so one CTA walks one utterance and the
// cost of a step is a latency chain.  The general kernel in post.cu pays an L2 round
// trip inside that chain (the rows "previous selection + 1" are only known after the
// previous step).  Here the chain never leaves the SM:
//
//   * The candidates of step s are idx[s] (known up front) and sel[s-1] + 1, and sel[s-1]
//     is a subset of the 8 candidates of step s-1.  So the 8 rows cand[s-1] + 1 are a
//     superset of what step s can need, and they are known one whole step early (as soon
//     as sel[s-2] is).  A producer warp fetches them speculatively with TMA bulk copies
//     (cp.async.bulk, one 4 KB row per instruction, completion on an mbarrier), together
//     with the idx[s] rows, the query row s and the log2-f0 of all twelve rows.
//   * Three generations of 13 rows live in shared memory: step s reads generation s
//     (its candidates) and generation s-1 (the four previously selected rows, which were
//     candidates there), while generation s+1 is in flight.  13 x 4 KB x 3 = 156 KB.
//   * Eight compute warps split the feature dimension: a thread keeps all 8 candidates x
//     {c.c, src.c, prev_j.c} = 48 partial dot products for its 4 columns, so each staged
//     row is read from shared memory once per step (52 KB) instead of once per candidate.
//     Products are fp32 (packed FFMA2); the 48 sums are reduced as a TREE — in-thread,
//     then a halving exchange over the warp (48 values -> 3 per lane in 45 shuffles), fp32,
//     depth 7, worst-case relative error < 1e-6 — and across warps in fp64.
//   * Warp 0 finishes the step: cosine distances 1 - x.c/(|x||c|) from the dot products and
//     carried reciprocal norms (one rsqrt per candidate; no sqrt/divide chain), the
//     reference's threshold edits, lower median and rank; it publishes the selection and
//     the producer picks it up and issues generation s+2.
//   The general kernel keeps the reference's direct-form distances in fp64; the
//   two agree to ~1e-6 in every cost, far inside the 1e-5 tie criterion.
//
// Traffic per frame: 13 rows = 53 KB against 9 rows = 36.9 KB algorithmic (the price of
// speculation); arithmetic per frame: 8 x 5 x 1024 fused difference-squares.
#include "common.cuh"
#include "kernels.cuh"

namespace knnsvc {

namespace {

constexpr int CS_K = 4;
constexpr int CS_C = 2 * CS_K;             // candidates per step
constexpr int CS_ACC = 6;                  // |c|^2, d2(src,c), d2(prev_j,c) j = 0..3
constexpr int CS_V = CS_C * CS_ACC;        // 48 partial sums per thread
constexpr int CS_WARPS = 8;                // compute warps
constexpr int CS_CT = CS_WARPS * 32;       // compute threads
constexpr int CS_THREADS = CS_CT + 32;     // + producer warp
constexpr int CS_ROWS = 13;                // rows per generation: 4 idx rows, 8 speculative rows, query row
constexpr int CS_GENS = 3;
constexpr int CS_MAX_DIM = 1024;

struct CsMeta {
  int64_t idx_g[CS_K];      // pool rows of idx[s]
  int64_t spec_g[CS_C];     // pool rows cand[s-1] + 1 (clamped)
  double lf0_idx[CS_K];     // log2(f0 + 1e-5) of those rows
  double lf0_spec[CS_C];
  double base, inv_src, lsrc; // 2*dist(src[s-1], src[s]), 1/|src[s]|, log2(src_f0[s] + 1e-5)
};

struct CsShared {
  CsMeta meta[CS_GENS];
  float part[CS_WARPS][CS_V];
  double prev_n2[CS_K];
  double prev_inv[CS_K];   // 1/|row| of the previous selections, carried from the step that scored them
  unsigned long long full_bar[CS_GENS];
  int sp[2][CS_K];          // candidate slot (0..7) of each selection, by step parity
  int prow[CS_K];           // row (0..11) of the previous generation holding each selected row
  volatile int sel_count;   // selections published so far
};

__device__ __forceinline__ uint32_t cs_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cs_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void cs_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cs_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// one row, global -> shared, completion credited to `bar`
__device__ __forceinline__ void cs_bulk_row(uint32_t dst, const float* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void cs_compute_sync() { asm volatile("bar.sync 1, %0;" ::"n"(CS_CT) : "memory"); }

__device__ __forceinline__ float2 lo2(const float4& v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(const float4& v) { return make_float2(v.z, v.w); }

// one halving step of the warp reduction: lanes with `upper` keep v[H..2H), the others v[0..H)
template <int H>
__device__ __forceinline__ void cs_halve(float (&v)[CS_V], bool upper, int xor_mask) {
#pragma unroll
  for (int k = 0; k < H; ++k) {
    const float keep = upper ? v[k + H] : v[k];
    const float send = upper ? v[k] : v[k + H];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, xor_mask);
  }
}

// ---- the arithmetic of one step, shared by the one-CTA and the cluster kernel.  Every operation is spelled
// out (explicit fma / mul / add intrinsics: no contraction left to the compiler), so the two kernels give
// the same bits for the same sums.
// fp64 sum of the eight per-slice partial sums p[0], p[stride], .. (slice = 128 feature dimensions: a
// compute warp of the one-CTA kernel, a CTA of the cluster kernel), as a fixed tree
__device__ __forceinline__ double cs_tree8(const float* p, int stride) {
  double v[CS_WARPS];
#pragma unroll
  for (int w = 0; w < CS_WARPS; ++w) v[w] = (double)p[w * stride];
  return __dadd_rn(__dadd_rn(__dadd_rn(v[0], v[4]), __dadd_rn(v[2], v[6])),
                   __dadd_rn(__dadd_rn(v[1], v[5]), __dadd_rn(v[3], v[7])));
}
// cosine distance 1 - a.b/(|a||b|) (lib_ongaku_test.py:162-165) from the dot product and carried reciprocal norms
__device__ __forceinline__ double cs_cos_dist(double dot, double inv_a, double inv_b) {
  return __fma_rn(-dot, __dmul_rn(inv_a, inv_b), 1.0);
}
// the reference's threshold edits of a concatenation cost (lib_ongaku_test.py:318-335)
__device__ __forceinline__ double cs_edit(double cc, double base, bool use_f0) {
  if (use_f0) return (base < 0.08 && cc < __dmul_rn(5.0, base)) ? 0.0 : cc;
  return cc > base ? __fma_rn(1.5, cc, -base) : cc;
}
// lower median of 4 = second smallest (torch.median, lib_ongaku_test.py:337,342)
// (plain compare-and-select: the costs are finite, and fmin / fmax spend half of their instructions on NaNs)
__device__ __forceinline__ double cs_median4(double c0, double c1, double c2, double c3) {
  const bool s01 = c0 < c1, s23 = c2 < c3;
  const double lo01 = s01 ? c0 : c1, hi01 = s01 ? c1 : c0;
  const double lo23 = s23 ? c2 : c3, hi23 = s23 ? c3 : c2;
  const double a = lo01 < lo23 ? lo23 : lo01;      // max of the two minima
  const double b = hi01 < hi23 ? hi01 : hi23;      // min of the two maxima
  return a < b ? a : b;
}
__device__ __forceinline__ double cs_total(double w, double med, double match, bool use_f0, double lcand,
                                           double lsrc) {
  const double t = __fma_rn(w, med, match);
  return use_f0 ? __dadd_rn(t, fabs(__dsub_rn(lcand, lsrc))) : t;
}
// one float4 column of a row against itself / another row: (x*x' then z*z') + (y*y' then w*w') as two fma chains
__device__ __forceinline__ float cs_col_dot(const float4& a, const float4& b) {
  float2 acc = make_float2(0.f, 0.f);
  acc = __ffma2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y), acc);
  acc = __ffma2_rn(make_float2(a.z, a.w), make_float2(b.z, b.w), acc);
  return acc.x + acc.y;
}

#ifdef KNNSVC_K5_PROFILE
__device__ long long g_k5_prof[8];
#define K5_T(i) do { const long long _t = clock64(); prof[i] += _t - t_last; t_last = _t; } while (0)
__device__ long long g_k5c_prof[16];
#define K5C_BEGIN() long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long t_last = clock64()
#define K5C_T(i) do { const long long _t = clock64(); prof[i] += _t - t_last; t_last = _t; } while (0)
#define K5C_END(base, cnt, cond) do { if (cond) for (int _i = 0; _i < (cnt); ++_i) g_k5c_prof[(base) + _i] = prof[_i]; } while (0)
#else
#define K5_T(i) do { } while (0)
#define K5C_BEGIN() do { } while (0)
#define K5C_T(i) do { } while (0)
#define K5C_END(base, cnt, cond) do { } while (0)
#endif

}  // namespace

__global__ void __launch_bounds__(CS_THREADS, 1) concat_cost_staged_kernel(
    const int64_t* __restrict__ idx, const float* __restrict__ src, const __grid_constant__ RowTable pool,
    int dim, const float* __restrict__ src_f0, const float* __restrict__ pool_f0, float concat_weight,
    const int64_t* __restrict__ utt_offsets, const double* __restrict__ base_all, const double* __restrict__ src_n2,
    int64_t* __restrict__ out_idx) {
  extern __shared__ __align__(128) unsigned char cs_raw[];
  float* rows = reinterpret_cast<float*>(cs_raw);                                   // [gen][row][dim]
  CsShared& sh = *reinterpret_cast<CsShared*>(cs_raw + (size_t)CS_GENS * CS_ROWS * dim * sizeof(float));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t f_begin = utt_offsets[blockIdx.x], f_end = utt_offsets[blockIdx.x + 1];
  const int64_t n = f_end - f_begin;
  if (n <= 0) return;
  const bool use_f0 = src_f0 != nullptr;
  const int64_t n_pool = pool.lo[pool.n];
  const uint32_t row_bytes = (uint32_t)dim * 4u;
  auto row_ptr = [&](int gen, int r) { return rows + ((size_t)gen * CS_ROWS + r) * dim; };

  if (tid == 0) {
    for (int g = 0; g < CS_GENS; ++g) cs_mbar_init(cs_smem_u32(&sh.full_bar[g]), 1);
    for (int j = 0; j < CS_K; ++j) {
      sh.sp[0][j] = j;      // "selection 0" = idx[0] itself, sitting in slots 0..3 of generation 0
      sh.sp[1][j] = j;
      sh.prow[j] = j;
    }
    sh.sel_count = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if (warp == CS_WARPS) {
    // ============================ producer warp ============================
    // generation 0: the four rows of idx[0] (the initial "previous selection")
    if (lane < CS_K) {
      const int64_t id = idx[f_begin * CS_K + lane];
      sh.meta[0].idx_g[lane] = id;
      cs_bulk_row(cs_smem_u32(row_ptr(0, lane)), table_row(pool, id, dim), row_bytes, cs_smem_u32(&sh.full_bar[0]));
    }
    __syncwarp();
    if (lane == 0) cs_mbar_expect_tx(cs_smem_u32(&sh.full_bar[0]), CS_K * row_bytes);
    // Everything addressed by the frame number alone is fetched one generation ahead into
    // registers, so the only latency between a published selection and the generation it
    // unlocks is the TMA copy itself (plus pool_f0 of the speculative rows).
    int64_t next_idx = 0;                       // lanes 8..11: idx[s]
    double next_base = 0.0, next_n2 = 1.0;      // lane 12: per-frame scalars of the query row
    float next_f0 = 0.f;
    if (n > 1) {
      if (lane >= CS_C && lane < CS_C + CS_K) next_idx = idx[(f_begin + 1) * CS_K + (lane - CS_C)];
      if (lane == CS_C + CS_K) {
        next_base = base_all[f_begin + 1];
        next_n2 = src_n2[f_begin + 1];
        if (use_f0) next_f0 = __ldg(src_f0 + f_begin + 1);
      }
    }
    for (int64_t s = 1; s < n; ++s) {
      const int g = (int)(s % CS_GENS), gp = (int)((s - 1) % CS_GENS);
      if (s >= 2) {  // cand[s-1] needs selection s-2
        if (lane == 0)
          while (sh.sel_count < (int)(s - 2)) {
          }
        __syncwarp();
        __threadfence_block();
      }
      const uint32_t bar = cs_smem_u32(&sh.full_bar[g]);
      int64_t id = -1;
      if (lane < CS_C) {                       // speculative rows: cand[s-1][lane] + 1
        int64_t c;
        if (s == 1) c = sh.meta[0].idx_g[lane & 3];
        else if (lane < CS_K) c = sh.meta[gp].idx_g[lane];
        else c = sh.meta[gp].spec_g[sh.sp[(s - 2) & 1][lane - CS_K]];
        id = c + 1 >= n_pool ? n_pool - 1 : c + 1;   // lib_ongaku_test.py:294-295
        sh.meta[g].spec_g[lane] = id;
        cs_bulk_row(cs_smem_u32(row_ptr(g, CS_K + lane)), table_row(pool, id, dim), row_bytes, bar);
      } else if (lane < CS_C + CS_K) {         // rows of idx[s]
        id = next_idx;
        sh.meta[g].idx_g[lane - CS_C] = id;
        cs_bulk_row(cs_smem_u32(row_ptr(g, lane - CS_C)), table_row(pool, id, dim), row_bytes, bar);
        if (s + 1 < n) next_idx = idx[(f_begin + s + 1) * CS_K + (lane - CS_C)];
      } else if (lane == CS_C + CS_K) {        // query row s and its per-frame scalars
        cs_bulk_row(cs_smem_u32(row_ptr(g, CS_ROWS - 1)), src + (f_begin + s) * dim, row_bytes, bar);
        sh.meta[g].base = next_base;
        sh.meta[g].inv_src = rsqrt(next_n2);
        sh.meta[g].lsrc = use_f0 ? log2((double)next_f0 + 1e-5) : 0.0;
        if (s + 1 < n) {
          next_base = base_all[f_begin + s + 1];
          next_n2 = src_n2[f_begin + s + 1];
          if (use_f0) next_f0 = __ldg(src_f0 + f_begin + s + 1);
        }
      }
      if (use_f0 && lane < CS_C + CS_K) {
        const double lf = log2((double)__ldg(pool_f0 + id) + 1e-5);
        if (lane < CS_C) sh.meta[g].lf0_spec[lane] = lf;
        else sh.meta[g].lf0_idx[lane - CS_C] = lf;
      }
      __syncwarp();
      if (lane == 0) cs_mbar_expect_tx(bar, CS_ROWS * row_bytes);   // release: publishes meta[g] too
    }
    return;
  }

  // ============================ compute warps ============================
  const int n4 = dim / 4;
  cs_mbar_wait(cs_smem_u32(&sh.full_bar[0]), 0);
  if (warp < CS_K) {  // |row|^2 of the four initial selections, one warp each
    const float4* r4 = reinterpret_cast<const float4*>(row_ptr(0, warp));
    double acc = 0.0;
    for (int c = lane; c < n4; c += 32) {
      const float4 v = r4[c];
      float t = v.x * v.x;
      t = fmaf(v.y, v.y, t);
      t = fmaf(v.z, v.z, t);
      t = fmaf(v.w, v.w, t);
      acc += (double)t;
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      sh.prev_n2[warp] = acc;
      sh.prev_inv[warp] = rsqrt(acc);
    }
    if (tid < CS_K) out_idx[f_begin * CS_K + tid] = sh.meta[0].idx_g[tid];
  }
  cs_compute_sync();

  double w_sticky = (double)concat_weight;   // meaningful in warp 0
#ifdef KNNSVC_K5_PROFILE
  long long prof[6] = {0, 0, 0, 0, 0, 0};
  long long t_last = clock64();
#endif
  for (int64_t s = 1; s < n; ++s) {
    const int g = (int)(s % CS_GENS), gp = (int)((s - 1) % CS_GENS);
    cs_mbar_wait(cs_smem_u32(&sh.full_bar[g]), (uint32_t)((s / CS_GENS) & 1));
    K5_T(0);
    const int* spv = sh.sp[(s - 1) & 1];
    const float4* cand4[CS_C];
#pragma unroll
    for (int m = 0; m < CS_C; ++m)
      cand4[m] = reinterpret_cast<const float4*>(row_ptr(g, m < CS_K ? m : CS_K + spv[m - CS_K]));
    const float4* src4 = reinterpret_cast<const float4*>(row_ptr(g, CS_ROWS - 1));
    const float4* prev4[CS_K];
#pragma unroll
    for (int j = 0; j < CS_K; ++j) prev4[j] = reinterpret_cast<const float4*>(row_ptr(gp, sh.prow[j]));

    // per-thread partial dot products over this thread's columns: acc[a][m], a = 0: c.c, 1: src.c, 2+j: prev_j.c
    float2 acc[CS_ACC][CS_C];
#pragma unroll
    for (int a = 0; a < CS_ACC; ++a)
#pragma unroll
      for (int m = 0; m < CS_C; ++m) acc[a][m] = make_float2(0.f, 0.f);
    for (int c = tid; c < n4; c += CS_CT) {
      float4 o[1 + CS_K];
      o[0] = src4[c];
#pragma unroll
      for (int j = 0; j < CS_K; ++j) o[1 + j] = prev4[j][c];
#pragma unroll
      for (int m = 0; m < CS_C; ++m) {
        const float4 cv = cand4[m][c];
        const float2 cl = lo2(cv), chh = hi2(cv);
        acc[0][m] = __ffma2_rn(cl, cl, acc[0][m]);
        acc[0][m] = __ffma2_rn(chh, chh, acc[0][m]);
#pragma unroll
        for (int a = 0; a < 1 + CS_K; ++a) {
          acc[1 + a][m] = __ffma2_rn(lo2(o[a]), cl, acc[1 + a][m]);
          acc[1 + a][m] = __ffma2_rn(hi2(o[a]), chh, acc[1 + a][m]);
        }
      }
    }
    K5_T(1);
    // warp tree reduction in fp32 (depth 2 + 5: worst-case relative error ~1e-6 / sqrt-ish typical 1e-7),
    // halving exchange: 48 -> 24 -> 12 -> 6 -> 3 values per lane, then one butterfly step
    float v[CS_V];
#pragma unroll
    for (int a = 0; a < CS_ACC; ++a)
#pragma unroll
      for (int m = 0; m < CS_C; ++m) v[a * CS_C + m] = acc[a][m].x + acc[a][m].y;
    cs_halve<24>(v, lane & 16, 16);
    cs_halve<12>(v, lane & 8, 8);
    cs_halve<6>(v, lane & 4, 4);
    cs_halve<3>(v, lane & 2, 2);
#pragma unroll
    for (int k = 0; k < 3; ++k) v[k] += __shfl_xor_sync(0xffffffffu, v[k], 1);
    if (!(lane & 1)) {
      const int off = ((lane & 16) ? 24 : 0) + ((lane & 8) ? 12 : 0) + ((lane & 4) ? 6 : 0) + ((lane & 2) ? 3 : 0);
#pragma unroll
      for (int k = 0; k < 3; ++k) sh.part[warp][off + k] = v[k];
    }
    K5_T(2);
    cs_compute_sync();
    K5_T(3);

    if (warp == 0) {
      const CsMeta& mt = sh.meta[g];
      const double base = mt.base;
      if (use_f0 && !(base < 0.08)) w_sticky = 0.0;  // sticky: persists for all later frames (lib_ongaku_test.py:332)
      // cross-warp sums in fp64: lane l owns flat entries l and l + 32 (flat = a * 8 + candidate)
      const double t0 = cs_tree8(&sh.part[0][lane], CS_V);
      const double t1 = cs_tree8(&sh.part[0][lane < CS_V - 32 ? lane + 32 : lane], CS_V);
      double q[CS_ACC];   // lane m < 8: the six sums of candidate m
#pragma unroll
      for (int a = 0; a < 4; ++a) q[a] = __shfl_sync(0xffffffffu, t0, a * CS_C + (lane & 7));
#pragma unroll
      for (int a = 4; a < CS_ACC; ++a) q[a] = __shfl_sync(0xffffffffu, t1, (a - 4) * CS_C + (lane & 7));
      double total = INFINITY, my_n2 = 0.0, my_inv = 0.0;
      int64_t my_id = 0;
      int my_row = 0;
      if (lane < CS_C) {
        const int slot = lane < CS_K ? 0 : spv[lane - CS_K];
        my_id = lane < CS_K ? mt.idx_g[lane] : mt.spec_g[slot];
        my_row = lane < CS_K ? lane : CS_K + slot;
        const double lcand = lane < CS_K ? mt.lf0_idx[lane] : mt.lf0_spec[slot];
        my_n2 = q[0];
        my_inv = rsqrt(my_n2);
        const double match = cs_cos_dist(q[1], mt.inv_src, my_inv);
        double cc[CS_K];
#pragma unroll
        for (int j = 0; j < CS_K; ++j) cc[j] = cs_edit(cs_cos_dist(q[2 + j], sh.prev_inv[j], my_inv), base, use_f0);
        total = cs_total(w_sticky, cs_median4(cc[0], cc[1], cc[2], cc[3]), match, use_f0, lcand, mt.lsrc);
      }
      int rank = 0;
#pragma unroll
      for (int j = 0; j < CS_C; ++j) {
        const double tj = __shfl_sync(0xffffffffu, total, j);
        rank += (tj < total) || (tj == total && j < lane);
      }
      __syncwarp();  // every lane has read prev_inv before it is overwritten
      if (lane < CS_C && rank < CS_K) {
        out_idx[(f_begin + s) * CS_K + rank] = my_id;
        sh.sp[s & 1][rank] = lane;
        sh.prow[rank] = my_row;
        sh.prev_n2[rank] = my_n2;
        sh.prev_inv[rank] = my_inv;
      }
      __syncwarp();
      if (lane == 0) {
        __threadfence_block();
        sh.sel_count = (int)s;   // the producer may now form cand[s] and issue generation s+1
      }
    }
    K5_T(4);
    cs_compute_sync();
    K5_T(5);
  }
#ifdef KNNSVC_K5_PROFILE
  if (tid == 0 && blockIdx.x == 0)
    for (int i = 0; i < 6; ++i) g_k5_prof[i] = prof[i];
#endif
}

// ------------------------------------------------------------------------------------------------
// Cluster variant: ONE UTTERANCE PER CLUSTER OF 8 CTAs, each CTA owning 128 of the feature dimensions.
//
// Measured on the one-CTA kernel (profiles/r2_k5_chain_experiment.txt): a step is bound by moving 13 rows =
// 53 KB into one SM and by one warp finishing all eight candidates.  A batch of utterances does not have to
// hide either — 148 CTAs run side by side — but a single long utterance (BASELINE cfg 2: 3001 frames)
// waits for every step.  Here
//   * CTA r holds only columns [128 r, 128 r + 128) of the 13 rows (6.5 KB per step, TMA bulk copies, the
//     same three-generation speculative ring as above);
//   * compute warp m scores CANDIDATE m on the CTA's slice (one float4 column per lane, the six sums
//     c.c, src.c, prev_j.c, a 5-level warp tree) and sends the six partial sums to warp m of all 8 CTAs with
//     st.async (remote shared-memory stores that complete a transaction count on the receiver's mbarrier:
//     no cluster barrier inside the loop);
//   * warp m of EVERY CTA then adds the 8 slices in fp64 and finishes candidate m's cost (4 lanes = the 4
//     previous selections, median over shuffles), posts it to its CTA, and every warp that needs the
//     selection ranks the eight costs for itself — no decision warp on the chain.  All 8 CTAs take the same
//     decisions from the same numbers, so nothing else crosses the cluster;
//   * a lone warp runs ~4.5 cycles per instruction, so the chain is kept free of everything that is not the
//     recurrence (profiles/r2b_k5_cluster.txt: with the fetch issued by the warps on the chain a step took
//     3000-3100 cycles, 2200 of them per-lane cp.async.bulk issue in the first version).  The rows are
//     fetched by warps of their own: two producer warps run two steps ahead with what the frame number
//     alone addresses (rows of idx[t], the query row and its scalars; the rows idx[t-1] + 1), four fetch
//     warps each rank the costs too and fetch the row that follows selection r, and an output warp writes
//     the result (rank 0) and tells the producers which ring buffer is free again.
// Arithmetic: slice r is exactly what compute warp r of the one-CTA kernel sums, the warp tree has the same
// levels (16, 8, 4, 2, 1), the cross-slice tree and the cost formulas are the shared functions above —
// the two kernels return the same bits.
constexpr int CL_C = 8;                          // CTAs per cluster (portable maximum)
constexpr int CL_SLICE = CS_MAX_DIM / CL_C;      // feature columns per CTA
constexpr int CL_W_PROD_A = CS_WARPS;            // rows of idx[t], query row t
constexpr int CL_W_PROD_B = CS_WARPS + 1;        // rows idx[t-1] + 1
constexpr int CL_W_FETCH = CS_WARPS + 2;         // .. + 5: row following selection r
constexpr int CL_W_OUT = CS_WARPS + 2 + CS_K;    // output
constexpr int CL_THREADS = (CL_W_OUT + 1) * 32;  // 15 warps
constexpr int CL_FULL_ARRIVALS = 2 + CS_K;       // per generation: producers A and B, four fetch warps
constexpr int CL_COST_ARRIVALS = CS_C + 1 + CS_K;   // per step: 8 costs; output warp and fetch warps "done with the step before"
static_assert(CL_C == CS_WARPS, "slice r of the cluster kernel = compute warp r of the one-CTA kernel");
static_assert(CL_SLICE == 128, "one float4 column per lane");

struct ClShared {
  float rows[CS_GENS][CS_ROWS][CL_SLICE];
  float xch[2][CS_C][CL_C][8];     // [step parity][candidate][source CTA]{c.c, src.c, prev0.c, -, prev1.c, prev2.c, prev3.c, -}
  CsMeta meta[CS_GENS];
  alignas(16) double cost[2][CS_C];   // [step parity]: total cost of each candidate (read as double2)
  double cinv[2][CS_C];            //   1/|candidate row|
  int64_t cid[2][CS_C];            //   its pool row
  int crow[2][CS_C];               //   its row (0..11) in the generation's ring buffer
  double init_inv[CS_K];           // 1/|row| of idx[0]
  unsigned long long full_bar[CS_GENS];    // CL_FULL_ARRIVALS + bytes: the generation has landed
  unsigned long long empty_bar[CS_GENS];   // output warp -> producers: the buffer's generation has been consumed
  unsigned long long xbar[2][CS_C];        // [step parity][candidate]: the 8 slices' partial sums have landed
  unsigned long long cbar;                 // CL_COST_ARRIVALS: the eight costs of a step are posted
};

__device__ __forceinline__ uint32_t cl_mapa(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void cl_st_async4(uint32_t raddr, float a, float b, float c, float d, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
                   raddr),
               "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(__float_as_uint(d)),
               "r"(rbar)
               : "memory");
}
__device__ __forceinline__ void cl_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// wait for a phase completed by peers' st.async (their complete_tx is a release at cluster scope)
__device__ __forceinline__ void cl_mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void cl_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// The selection of a step from its eight posted costs, by a whole warp: lane l compares candidate l & 7 with
// candidates 2 (l >> 3) and 2 (l >> 3) + 1, two shuffle-adds give every lane the rank of candidate l & 7
// (ties to the lower candidate slot), and ONE warp-wide OR gathers the four selected candidates: byte r of the
// result = (candidate slot << 4) | `tag` of the candidate of rank r (tag: 4 bits of the caller's, lanes 0..7).
__device__ __forceinline__ unsigned cl_select(const double* cost, int lane, int tag) {
  const int a = lane & 7, b = (lane >> 3) * 2;
  const double t = cost[a];
  const double2 u = *reinterpret_cast<const double2*>(cost + b);
  int cnt = (int)((u.x < t) || (u.x == t && b < a)) + (int)((u.y < t) || (u.y == t && b + 1 < a));
  cnt += __shfl_xor_sync(0xffffffffu, cnt, 8);
  cnt += __shfl_xor_sync(0xffffffffu, cnt, 16);
  const unsigned mine = (lane < CS_C && cnt < CS_K) ? (unsigned)((a << 4) | (tag & 15)) << (8 * cnt) : 0u;
  return __reduce_or_sync(0xffffffffu, mine);
}
__device__ __forceinline__ int cl_sel_slot(unsigned packed, int r) { return (int)(packed >> (8 * r + 4)) & 7; }
__device__ __forceinline__ int cl_sel_tag(unsigned packed, int r) { return (int)(packed >> (8 * r)) & 15; }

__global__ void __cluster_dims__(CL_C, 1, 1) __launch_bounds__(CL_THREADS, 1) concat_cost_cluster_kernel(
    const int64_t* __restrict__ idx, const float* __restrict__ src, const __grid_constant__ RowTable pool,
    int dim, const float* __restrict__ src_f0, const float* __restrict__ pool_f0,
    const double* __restrict__ lf0_tab, float concat_weight,
    const int64_t* __restrict__ utt_offsets, const double* __restrict__ base_all, const double* __restrict__ src_n2,
    int64_t* __restrict__ out_idx) {
  __shared__ __align__(128) ClShared sh;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int utt = blockIdx.x / CL_C;
  const int64_t f_begin = utt_offsets[utt], f_end = utt_offsets[utt + 1];
  if (f_end - f_begin <= 0) return;                    // (the same decision in all CTAs of the cluster)
  const int n = (int)(f_end - f_begin);                // (the launcher admits utterances below 2^31 frames)
  const bool use_f0 = src_f0 != nullptr;
  // log2 f0 of the pool rows: from the per-call table when there is one (the compute warps load their candidate's
  // value themselves), else computed by the warp that fetches the row
  const bool row_lf0 = use_f0 && lf0_tab == nullptr;
  const int64_t n_pool = pool.lo[pool.n];
  const int s_lo = (int)rank * CL_SLICE;               // first feature column of this CTA
  const int slice_len = dim - s_lo < 0 ? 0 : (dim - s_lo > CL_SLICE ? CL_SLICE : dim - s_lo);
  const uint32_t slice_bytes = (uint32_t)slice_len * 4u;
  auto clamp_next = [&](int64_t c) { return c + 1 >= n_pool ? n_pool - 1 : c + 1; };   // lib_ongaku_test.py:294-295

  if (tid == 0) {
    for (int g = 0; g < CS_GENS; ++g) {
      cs_mbar_init(cs_smem_u32(&sh.full_bar[g]), CL_FULL_ARRIVALS);
      cs_mbar_init(cs_smem_u32(&sh.empty_bar[g]), 1);
    }
    for (int p = 0; p < 2; ++p)
      for (int m = 0; m < CS_C; ++m) cs_mbar_init(cs_smem_u32(&sh.xbar[p][m]), 1);
    cs_mbar_init(cs_smem_u32(&sh.cbar), CL_COST_ARRIVALS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  cl_cluster_sync();      // every CTA's barriers are live before a peer sends to them

  if (warp == CL_W_PROD_A) {
    // ============ producer A: rows of idx[t] (lanes 0..3), query row t and its scalars (lane 4) ============
    // Global operands of generation t + 1 are loaded into registers while generation t is issued.
    if (lane < CS_K) {      // generation 0: this CTA's slice of the four rows of idx[0] (no other rows)
      const int64_t id = idx[f_begin * CS_K + lane];
      sh.meta[0].idx_g[lane] = id;
      if (slice_bytes)
        cs_bulk_row(cs_smem_u32(&sh.rows[0][lane][0]), table_row(pool, id, dim) + s_lo, slice_bytes,
                    cs_smem_u32(&sh.full_bar[0]));
    }
    __syncwarp();
    if (lane == 0) cs_mbar_expect_tx(cs_smem_u32(&sh.full_bar[0]), CS_K * slice_bytes);
    int64_t next_idx = 0;
    double next_base = 0.0, next_n2 = 1.0;
    float next_f0 = 0.f, next_pf0 = 0.f;
    auto prefetch = [&](int t) {            // operands of generation t -> registers
      if (t >= n) return;
      if (lane < CS_K) {
        next_idx = idx[(f_begin + t) * CS_K + lane];
        if (row_lf0) next_pf0 = __ldg(pool_f0 + next_idx);
      } else if (lane == CS_K) {
        next_base = base_all[f_begin + t];
        next_n2 = src_n2[f_begin + t];
        if (use_f0) next_f0 = __ldg(src_f0 + f_begin + t);
      }
    };
    prefetch(1);
    for (int t = 1; t < n; ++t) {
      const int g = t % CS_GENS;
      const int64_t id = next_idx;
      const double b = next_base, n2 = next_n2;
      const float f0s = next_f0, f0p = next_pf0;
      prefetch(t + 1);
      // the ring buffer is free once the step that read generation t-3 as "previous" (step t-2) is decided
      if (t >= CS_GENS) cs_mbar_wait(cs_smem_u32(&sh.empty_bar[g]), (uint32_t)((t / CS_GENS - 1) & 1));
      const uint32_t bar = cs_smem_u32(&sh.full_bar[g]);
      if (lane < CS_K) {
        sh.meta[g].idx_g[lane] = id;
        if (slice_bytes)
          cs_bulk_row(cs_smem_u32(&sh.rows[g][lane][0]), table_row(pool, id, dim) + s_lo, slice_bytes, bar);
        if (row_lf0) sh.meta[g].lf0_idx[lane] = log2((double)f0p + 1e-5);
      } else if (lane == CS_K) {
        if (slice_bytes)
          cs_bulk_row(cs_smem_u32(&sh.rows[g][CS_ROWS - 1][0]), src + (f_begin + t) * dim + s_lo, slice_bytes, bar);
        sh.meta[g].base = b;
        sh.meta[g].inv_src = rsqrt(n2);
        sh.meta[g].lsrc = use_f0 ? log2((double)f0s + 1e-5) : 0.0;
      }
      __syncwarp();
      if (lane == 0) cs_mbar_expect_tx(bar, (CS_K + 1) * slice_bytes);   // release: publishes its part of meta[g]
    }
  } else if (warp == CL_W_PROD_B) {
    // ============ producer B: speculative rows 0..3 of generation t = the rows after idx[t-1] ============
    // (candidates 0..3 of step t-1 are idx[t-1] whatever is selected; "step 0" has idx[0] in slots 0..3)
    if (lane == 0) cl_mbar_arrive(cs_smem_u32(&sh.full_bar[0]));
    int64_t next_id = 0;
    float next_pf0 = 0.f;
    auto prefetch = [&](int t) {
      if (t >= n || lane >= CS_K) return;
      next_id = clamp_next(idx[(f_begin + t - 1) * CS_K + lane]);
      if (row_lf0) next_pf0 = __ldg(pool_f0 + next_id);
    };
    prefetch(1);
    for (int t = 1; t < n; ++t) {
      const int g = t % CS_GENS;
      const int64_t id = next_id;
      const float f0p = next_pf0;
      prefetch(t + 1);
      if (t >= CS_GENS) cs_mbar_wait(cs_smem_u32(&sh.empty_bar[g]), (uint32_t)((t / CS_GENS - 1) & 1));
      const uint32_t bar = cs_smem_u32(&sh.full_bar[g]);
      if (lane < CS_K) {
        sh.meta[g].spec_g[lane] = id;
        if (slice_bytes)
          cs_bulk_row(cs_smem_u32(&sh.rows[g][CS_K + lane][0]), table_row(pool, id, dim) + s_lo, slice_bytes, bar);
        if (row_lf0) sh.meta[g].lf0_spec[lane] = log2((double)f0p + 1e-5);
      }
      __syncwarp();
      if (lane == 0) cs_mbar_expect_tx(bar, CS_K * slice_bytes);
    }
  } else if (warp >= CL_W_FETCH && warp < CL_W_FETCH + CS_K) {
    // ============ fetch warp r: speculative row 4 + r of generation s + 1 = the row after candidate 4 + r of
    // step s, which is the row after selection r of step s-1 ============
    // (generations 0 and 1 need none: step 1 follows the initial selection, whose next rows are producer B's)
    const int r = warp - CL_W_FETCH;
    const uint32_t cbar = cs_smem_u32(&sh.cbar);
    if (lane == 0) {
      cl_mbar_arrive(cs_smem_u32(&sh.full_bar[0]));
      if (n > 1) cl_mbar_arrive(cs_smem_u32(&sh.full_bar[1]));
    }
    for (int s = 1; s < n; ++s) {
      const int pp = (s & 1) ^ 1;
      int64_t sel_id;
      if (s >= 2) {
        cs_mbar_wait(cbar, (uint32_t)(s & 1));                            // costs of step s-1 (phase s-2)
        sel_id = sh.cid[pp][cl_sel_slot(cl_select(sh.cost[pp], lane, 0), r)];
      } else {
        sel_id = idx[f_begin * CS_K + r];                                 // selection 0 = idx[0]
      }
      __syncwarp();
      if (lane == 0) {
        cl_mbar_arrive(cbar);                                             // done with the costs of step s-1
        if (s + 1 < n) {
          const int g = (s + 1) % CS_GENS;
          const int64_t id = clamp_next(clamp_next(sel_id));
          const uint32_t bar = cs_smem_u32(&sh.full_bar[g]);
          if (slice_bytes)
            cs_bulk_row(cs_smem_u32(&sh.rows[g][2 * CS_K + r][0]), table_row(pool, id, dim) + s_lo, slice_bytes, bar);
          sh.meta[g].spec_g[CS_K + r] = id;
          if (row_lf0) sh.meta[g].lf0_spec[CS_K + r] = log2((double)__ldg(pool_f0 + id) + 1e-5);
          cs_mbar_expect_tx(bar, slice_bytes);                            // release
        }
      }
      __syncwarp();
    }
  } else if (warp == CL_W_OUT) {
    // ============ output warp (off the chain) ============
    if (rank == 0 && lane < CS_K) out_idx[f_begin * CS_K + lane] = idx[f_begin * CS_K + lane];
    for (int s = 1; s < n; ++s) {
      const int par = s & 1;
      if (lane == 0) cl_mbar_arrive(cs_smem_u32(&sh.cbar));             // done with the costs of step s-1
      cs_mbar_wait(cs_smem_u32(&sh.cbar), (uint32_t)((s - 1) & 1));     // the eight costs of step s
      const unsigned sel = cl_select(sh.cost[par], lane, 0);
      if (rank == 0 && lane < CS_K) out_idx[(f_begin + s) * CS_K + lane] = sh.cid[par][cl_sel_slot(sel, lane)];
      __syncwarp();
      // generation s-1 was read for the last time (as the previous one) by step s
      if (lane == 0) cl_mbar_arrive(cs_smem_u32(&sh.empty_bar[(s - 1) % CS_GENS]));
    }
  } else {
    // ============ compute warps: warp m scores candidate m ============
    const int m = warp;
    const bool has_col = 4 * lane < slice_len;
    const int j = lane & 3;
    if (warp < CS_K) {  // |row|^2 of the four initial selections (whole rows, from global memory: once)
      const int64_t id0 = idx[f_begin * CS_K + warp];
      const float4* r4 = reinterpret_cast<const float4*>(table_row(pool, id0, dim));
      const int n4 = dim / 4;
      double acc = 0.0;
      for (int c = lane; c < n4; c += 32) {
        const float4 v = __ldg(r4 + c);
        float t = v.x * v.x;
        t = fmaf(v.y, v.y, t);
        t = fmaf(v.z, v.z, t);
        t = fmaf(v.w, v.w, t);
        acc += (double)t;
      }
      acc = warp_sum(acc);
      if (lane == 0) sh.init_inv[warp] = rsqrt(acc);
    }
    cs_compute_sync();
    // the previous selection, in registers: candidate slot of each rank, its ring-buffer row, 1/|row| of rank j
    int sp0 = 0, sp1 = 1, sp2 = 2, sp3 = 3;
    int pr[CS_K] = {0, 1, 2, 3};
    double pinv = sh.init_inv[j];
    double w_sticky = (double)concat_weight;
    // loop invariants: where this lane's st.async goes (both step parities), which exchanged value it sums
    const uint32_t dst = (uint32_t)(lane & 7);
    const int xoff = (lane & 12) == 0 ? (j == 0 ? 2 : 3 + j) : ((lane & 12) == 4 ? 0 : 1);   // lanes 0..3 prev_j.c, 4..7 c.c, 8.. src.c
    const uint32_t lbar0 = cs_smem_u32(&sh.xbar[0][m]), lbar1 = cs_smem_u32(&sh.xbar[1][m]);
    const uint32_t rbar0 = cl_mapa(lbar0, dst), rbar1 = cl_mapa(lbar1, dst);
    const uint32_t rdat0 = cl_mapa(cs_smem_u32(&sh.xch[0][m][rank][(lane >> 4) * 4]), dst);
    const uint32_t rdat1 = cl_mapa(cs_smem_u32(&sh.xch[1][m][rank][(lane >> 4) * 4]), dst);
    const float* xsum0 = &sh.xch[0][m][0][xoff];
    const float* xsum1 = &sh.xch[1][m][0][xoff];
    const uint32_t cbar = cs_smem_u32(&sh.cbar);
    cs_mbar_wait(cs_smem_u32(&sh.full_bar[0]), 0);
    int g = 1, gp = 0;
    uint32_t full_phase = 1u;          // parity to wait for on each ring buffer's barrier (generation 0 is waited for here)
    K5C_BEGIN();
    for (int s = 1; s < n; ++s) {
      const int par = s & 1;
      // generation s (issued during step s-1) first: nothing of it may be read before its barrier is seen
      cs_mbar_wait(cs_smem_u32(&sh.full_bar[g]), (full_phase >> g) & 1u);
      full_phase ^= 1u << g;
      K5C_T(0);
      const uint32_t lbar = par ? lbar1 : lbar0;
      if (lane == 0) cs_mbar_expect_tx(lbar, CL_C * 32u);
      const float* rg = &sh.rows[g][0][4 * lane];
      const float* rgp = &sh.rows[gp][0][4 * lane];
      float4 sv = make_float4(0.f, 0.f, 0.f, 0.f), cv = sv;
      if (has_col) {      // what does not depend on selection s-1 is read first
        sv = *reinterpret_cast<const float4*>(rg + (CS_ROWS - 1) * CL_SLICE);
        if (m < CS_K) cv = *reinterpret_cast<const float4*>(rg + m * CL_SLICE);
      }
      if (s >= 2) {
        // selection s-1: every compute warp ranks the eight costs of step s-1 for itself
        cs_mbar_wait(cbar, (uint32_t)par);                              // phase s-2
        K5C_T(1);
        const int pp = par ^ 1;
        const double ci = sh.cinv[pp][lane & 7];
        const unsigned sel = cl_select(sh.cost[pp], lane, sh.crow[pp][lane & 7]);
        sp0 = cl_sel_slot(sel, 0), sp1 = cl_sel_slot(sel, 1), sp2 = cl_sel_slot(sel, 2), sp3 = cl_sel_slot(sel, 3);
#pragma unroll
        for (int jj = 0; jj < CS_K; ++jj) pr[jj] = cl_sel_tag(sel, jj);
        pinv = __shfl_sync(0xffffffffu, ci, cl_sel_slot(sel, j));
      }
      K5C_T(2);
      const int slot = m < CS_K ? 0 : (m == 4 ? sp0 : (m == 5 ? sp1 : (m == 6 ? sp2 : sp3)));
      const int crow = m < CS_K ? m : CS_K + slot;
      const CsMeta& mt = sh.meta[g];
      const int64_t c_my = m < CS_K ? mt.idx_g[m] : mt.spec_g[slot];      // the candidate's pool row
      double lcand = 0.0;
      if (use_f0 && !row_lf0) lcand = __ldg(lf0_tab + c_my);               // needed at the end of the step only
      // the six partial sums of candidate m over this CTA's slice: v[0] c.c, v[1] src.c, v[2 + jj] prev_jj.c
      float v[CS_ACC];
#pragma unroll
      for (int a = 0; a < CS_ACC; ++a) v[a] = 0.f;
      if (has_col) {
        if (m >= CS_K) cv = *reinterpret_cast<const float4*>(rg + crow * CL_SLICE);
        v[0] = cs_col_dot(cv, cv);
        v[1] = cs_col_dot(sv, cv);
#pragma unroll
        for (int jj = 0; jj < CS_K; ++jj)
          v[2 + jj] = cs_col_dot(*reinterpret_cast<const float4*>(rgp + pr[jj] * CL_SLICE), cv);
      }
      // warp tree, levels 16, 8, 4, 2, 1 as in the one-CTA kernel: lanes >= 16 end up with sums 3..5
      const bool upper = (lane & 16) != 0;
      float r[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float keep = upper ? v[k + 3] : v[k];
        const float send = upper ? v[k] : v[k + 3];
        r[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1)
#pragma unroll
        for (int k = 0; k < 3; ++k) r[k] += __shfl_xor_sync(0xffffffffu, r[k], o);
      // to warp m of every CTA of the cluster (this one included)
      if ((lane & 15) < CL_C) cl_st_async4(par ? rdat1 : rdat0, r[0], r[1], r[2], 0.f, par ? rbar1 : rbar0);
      // while the sums travel: what the cost needs besides them
      const double base = mt.base, inv_src = mt.inv_src;
      if (row_lf0) lcand = m < CS_K ? mt.lf0_idx[m] : mt.lf0_spec[slot];
      const double lsrc = mt.lsrc;
      if (use_f0 && !(base < 0.08)) w_sticky = 0.0;  // sticky: persists for all later frames (lib_ongaku_test.py:332)
      K5C_T(3);
      cl_mbar_wait_cluster(lbar, (uint32_t)(((s - 1) >> 1) & 1));
      K5C_T(4);
      // candidate m's cost: lanes 0..3 own one previous selection each, lanes 4..7 |c|^2, lanes 8..11 src.c
      const double mine = cs_tree8(par ? xsum1 : xsum0, 8);
      const double n2 = __shfl_sync(0xffffffffu, mine, 4);
      const double d_src = __shfl_sync(0xffffffffu, mine, 8);
      const double my_inv = rsqrt(n2);
      const double match = cs_cos_dist(d_src, inv_src, my_inv);
      const double cc = cs_edit(cs_cos_dist(mine, pinv, my_inv), base, use_f0);    // meaningful in lanes 0..3
      const double c0 = __shfl_sync(0xffffffffu, cc, 0), c1 = __shfl_sync(0xffffffffu, cc, 1);
      const double c2 = __shfl_sync(0xffffffffu, cc, 2), c3 = __shfl_sync(0xffffffffu, cc, 3);
      const double total = cs_total(w_sticky, cs_median4(c0, c1, c2, c3), match, use_f0, lcand, lsrc);
      if (lane == 0) {
        sh.cost[par][m] = total;
        sh.cinv[par][m] = my_inv;
        sh.cid[par][m] = c_my;
        sh.crow[par][m] = crow;
        cl_mbar_arrive(cbar);      // release
      }
      gp = g;
      g = g == CS_GENS - 1 ? 0 : g + 1;
      K5C_T(5);
    }
    K5C_END(0, 6, lane == 0 && warp == 0 && blockIdx.x == 0);
  }
  __syncwarp();
  cl_cluster_sync();      // nobody leaves while a peer could still address its shared memory
}


size_t concat_staged_smem_bytes(int dim) {
  return (size_t)CS_GENS * CS_ROWS * dim * sizeof(float) + sizeof(CsShared);
}

bool concat_staged_eligible(const float* src, const RowTable& pool, int dim) {
  bool ok = dim >= 4 && dim % 4 == 0 && dim <= CS_MAX_DIM && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
  for (int s = 0; s < pool.n; ++s) ok = ok && (reinterpret_cast<uintptr_t>(pool.base[s]) & 15) == 0;
  return ok;
}

// The cluster kernel is for a few long utterances: every utterance needs a cluster of 8 SMs of its own,
// resident at once (a second wave would wait for a whole utterance).
bool concat_cluster_fits(int n_utt) {
  static PerDevice cache;           // per device: 0 = not asked yet, else 1 + clusters the device hosts at once
  std::atomic<int>* slot = cache.slot();
  int n_clusters = slot ? slot->load(std::memory_order_relaxed) - 1 : -1;
  if (n_clusters < 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CL_C);
    cfg.blockDim = dim3(CL_THREADS);
    cfg.dynamicSmemBytes = 0;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL_C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int nc = 0;
    const cudaError_t e = cudaOccupancyMaxActiveClusters(&nc, concat_cost_cluster_kernel, &cfg);
    if (e != cudaSuccess) (void)cudaGetLastError();
    n_clusters = e == cudaSuccess ? nc : 0;
    if (slot) slot->store(n_clusters + 1, std::memory_order_relaxed);
  }
  return n_utt >= 1 && n_utt <= n_clusters;
}

int launch_concat_cost_cluster(const int64_t* idx, const float* src, const RowTable& pool, int dim,
                               const float* src_f0, const float* pool_f0, const double* lf0_tab, float concat_weight,
                               const int64_t* utt_offsets_dev, int n_utt, const double* base, const double* n2,
                               int64_t* out_idx, cudaStream_t stream) {
  concat_cost_cluster_kernel<<<n_utt * CL_C, CL_THREADS, 0, stream>>>(idx, src, pool, dim, src_f0, pool_f0, lf0_tab,
                                                                      concat_weight, utt_offsets_dev, base, n2, out_idx);
  KNN_LAUNCH_CHECK();
#ifdef KNNSVC_K5_PROFILE
  {  // debugging aid: cycles of cluster 0 / rank 0 per phase, accumulated over the launch
    long long h[16] = {0};
    cudaStreamSynchronize(stream);
    cudaMemcpyFromSymbol(h, g_k5c_prof, sizeof(h));
    fprintf(stderr, "[k5 cluster prof] warp0: gen-wait %lld cost-wait %lld rank %lld math+send+issue %lld xchg-wait %lld "
            "cost+publish %lld\n", h[0], h[1], h[2], h[3], h[4], h[5]);
  }
#endif
  return 0;
}

int launch_concat_cost_staged(const int64_t* idx, const float* src, const RowTable& pool, int dim,
                              const float* src_f0, const float* pool_f0, float concat_weight,
                              const int64_t* utt_offsets_dev, int n_utt, const double* base, const double* n2,
                              int64_t* out_idx, cudaStream_t stream) {
  const size_t smem = concat_staged_smem_bytes(dim);
  static PerDevice attr;
  KNN_SMEM_ATTR(attr, concat_cost_staged_kernel, smem);
  concat_cost_staged_kernel<<<n_utt, CS_THREADS, smem, stream>>>(idx, src, pool, dim, src_f0, pool_f0,
                                                                 concat_weight, utt_offsets_dev, base, n2, out_idx);
  KNN_LAUNCH_CHECK();
#ifdef KNNSVC_K5_PROFILE
  {  // debugging aid: cycles of block 0 / thread 0 per phase, accumulated over the launch
    long long h[8] = {0};
    cudaStreamSynchronize(stream);
    cudaMemcpyFromSymbol(h, g_k5_prof, sizeof(h));
    fprintf(stderr, "[k5 prof] wait %lld math %lld reduce %lld sync1 %lld final %lld sync2 %lld\n", h[0], h[1], h[2], h[3],
            h[4], h[5]);
    long long z[8] = {0};
    cudaMemcpyToSymbol(g_k5_prof, z, sizeof(z));
  }
#endif
  return 0;
}

}  // namespace knnsvc
