// Row pre-pass, full-matrix cosine distance and the exact brute-force kNN.
//
//  prepare_rows   : |x| and the unit-normalised fp16 operand of the tcgen05 filter
//                   (replaces torch.norm at lib_ongaku_test.py:150-151).
//  cosine_dist    : the [T,Np] matrix fast_cosine_dist returns
//                   (lib_ongaku_test.py:148-175, ddsp_matcher.py:213-221) — API parity
//                   only; the fused kNN never materialises it.
//  knn_exact      : CUDA-core exact kNN, fp64 accumulation.  It is the decision
//                   procedure for rows the tensor-core filter cannot decide
//                   (massive ties) and an independent GPU check in the tests.
#include <cuda_bf16.h>

#include "common.cuh"
#include "kernels.cuh"

namespace knnsvc {

// ------------------------------------------------------------------ prepare_rows
// One warp per row; HBM-bound: reads dim*4 B, writes dim_pad*2 + 4 B per row.
__global__ void __launch_bounds__(256) prepare_rows_kernel(
    const float* __restrict__ x, int64_t rows, int dim, int64_t ld,
    __half* __restrict__ hout, int dim_pad, double* __restrict__ norms, int* __restrict__ bad_rows, int bf16,
    float* __restrict__ max_err) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const bool vec = (dim % 4 == 0) && (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                   (dim_pad % 4 == 0);
  float warp_max2 = 0.f;   // largest squared rounding error (2^10-scaled) over this warp's rows
  for (int64_t r = warp; r < rows; r += nwarps) {
    const float* xr = x + r * ld;
    double ss = 0.0;
    if (vec) {
      const float4* x4 = reinterpret_cast<const float4*>(xr);
      for (int c = lane; c < dim / 4; c += 32) {
        float4 v = __ldg(x4 + c);
        ss += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
      }
    } else {
      for (int c = lane; c < dim; c += 32) {
        float v = __ldg(xr + c);
        ss += (double)v * v;
      }
    }
    ss = warp_sum(ss);
    const double nrm = sqrt(ss);
    const bool bad = !(nrm > 0.0) || !isfinite(nrm);
    if (lane == 0) {
      norms[r] = nrm;
      if (bad) atomicAdd(bad_rows, 1);
    }
    const float sc = bad ? 0.0f : (float)((double)kHalfScale / nrm);
    // rho^2 = sum (u_i - x_i/|x|)^2 with u_i = operand_i / 2^10: the row's actual rounding error.
    // In the 2^10-scaled domain that is (operand_i - v_i) with v_i = fl32(x_i * sc) the value that
    // was rounded (the subtraction is exact); v_i itself is within 2 * 2^-24 relative of the exact
    // x_i * 2^10 / |x| (fp32 product, fp32 sc) — covered by the margins applied at the end.
    float err2 = 0.f;
    __half* hr = hout + r * (int64_t)dim_pad;
    if (vec) {
      const float4* x4 = reinterpret_cast<const float4*>(xr);
      uint2* h4 = reinterpret_cast<uint2*>(hr);
      for (int c = lane; c < dim_pad / 4; c += 32) {
        float4 v = (c < dim / 4) ? __ldg(x4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        uint2 o;
        float2 ua, ub;
        if (bf16) {
          __nv_bfloat162 a = __floats2bfloat162_rn(v.x * sc, v.y * sc);
          __nv_bfloat162 b = __floats2bfloat162_rn(v.z * sc, v.w * sc);
          o.x = *reinterpret_cast<uint32_t*>(&a);
          o.y = *reinterpret_cast<uint32_t*>(&b);
          ua = __bfloat1622float2(a);
          ub = __bfloat1622float2(b);
        } else {
          __half2 a = __floats2half2_rn(v.x * sc, v.y * sc);
          __half2 b = __floats2half2_rn(v.z * sc, v.w * sc);
          o.x = *reinterpret_cast<uint32_t*>(&a);
          o.y = *reinterpret_cast<uint32_t*>(&b);
          ua = __half22float2(a);
          ub = __half22float2(b);
        }
        h4[c] = o;
        const float e0 = ua.x - v.x * sc, e1 = ua.y - v.y * sc, e2 = ub.x - v.z * sc, e3 = ub.y - v.w * sc;
        err2 = fmaf(e0, e0, fmaf(e1, e1, fmaf(e2, e2, fmaf(e3, e3, err2))));
      }
    } else {
      for (int c = lane; c < dim_pad; c += 32) {
        const float x0 = (c < dim) ? __ldg(xr + c) : 0.0f;
        const float v = x0 * sc;
        float u;
        if (bf16) {
          const __nv_bfloat16 b = __float2bfloat16_rn(v);
          reinterpret_cast<__nv_bfloat16*>(hr)[c] = b;
          u = __bfloat162float(b);
        } else {
          const __half h = __float2half_rn(v);
          hr[c] = h;
          u = __half2float(h);
        }
        const float e = u - v;
        err2 = fmaf(e, e, err2);
      }
    }
    if (max_err) {
      err2 = warp_sum(err2);
      if (!bad) warp_max2 = fmaxf(warp_max2, err2);
    }
  }
  // one atomic per warp (not per row: millions of same-address atomics serialise in L2).  Positive
  // floats order like their bit patterns; 1e-3 relative + 1e-6 absolute cover the fp32 evaluation.
  if (max_err && lane == 0 && warp_max2 > 0.f)
    atomicMax(reinterpret_cast<int*>(max_err), __float_as_int(sqrtf(warp_max2) * (1.001f / kHalfScale) + 1e-6f));
}

int launch_prepare_rows(const float* x, int64_t rows, int dim, int64_t ld, void* half_out, int dim_pad,
                        double* norms, int* bad_rows, float* max_err, cudaStream_t stream) {
  if (rows == 0) return 0;
  int64_t blocks = ceil_div64(rows, 8);
  if (blocks > 148 * 16) blocks = 148 * 16;
  prepare_rows_kernel<<<(unsigned)blocks, 256, 0, stream>>>(x, rows, dim, ld, (__half*)half_out, dim_pad,
                                                            norms, bad_rows, opt_bf16(), max_err);
  KNN_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------ cosine_dist (full matrix)
// 64x64 output tile, 16-wide k slabs, 256 threads with a 4x4 micro-tile each.
// Norms are recomputed from the same slabs so the kernel is self-contained.
constexpr int CD_T = 64, CD_K = 16;

__global__ void __launch_bounds__(256) cosine_dist_kernel(const float* __restrict__ q, int64_t nq,
                                                          const float* __restrict__ p, int64_t np_, int dim,
                                                          float* __restrict__ out) {
  __shared__ float sq[CD_K][CD_T + 1];
  __shared__ float sp[CD_K][CD_T + 1];
  __shared__ float nq2[CD_T], np2[CD_T];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t q0 = (int64_t)blockIdx.y * CD_T, p0 = (int64_t)blockIdx.x * CD_T;
  float acc[4][4] = {};
  float nacc = 0.f;  // threads 0..63 accumulate |q|^2, 64..127 |p|^2
  for (int k0 = 0; k0 < dim; k0 += CD_K) {
    for (int e = tid; e < CD_T * CD_K; e += 256) {
      int r = e / CD_K, c = e % CD_K;
      int64_t qr = q0 + r, pr = p0 + r;
      sq[c][r] = (qr < nq && k0 + c < dim) ? __ldg(q + qr * dim + k0 + c) : 0.f;
      sp[c][r] = (pr < np_ && k0 + c < dim) ? __ldg(p + pr * dim + k0 + c) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < CD_K; ++c) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sq[c][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sp[c][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (tid < 64) {
#pragma unroll
      for (int c = 0; c < CD_K; ++c) nacc = fmaf(sq[c][tid], sq[c][tid], nacc);
    } else if (tid < 128) {
#pragma unroll
      for (int c = 0; c < CD_K; ++c) nacc = fmaf(sp[c][tid - 64], sp[c][tid - 64], nacc);
    }
    __syncthreads();
  }
  if (tid < 64) nq2[tid] = nacc;
  else if (tid < 128) np2[tid - 64] = nacc;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t qr = q0 + ty * 4 + i;
    if (qr >= nq) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int64_t pr = p0 + tx * 4 + j;
      if (pr >= np_) continue;
      float den = sqrtf(nq2[ty * 4 + i]) * sqrtf(np2[tx * 4 + j]);
      out[qr * np_ + pr] = 1.0f - acc[i][j] / den;
    }
  }
}

int launch_cosine_dist(const float* q, int64_t nq, const float* p, int64_t np_, int dim, float* out,
                       cudaStream_t stream) {
  if (nq == 0 || np_ == 0) return 0;
  dim3 grid((unsigned)ceil_div64(np_, CD_T), (unsigned)ceil_div64(nq, CD_T));
  cosine_dist_kernel<<<grid, 256, 0, stream>>>(q, nq, p, np_, dim, out);
  KNN_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------ knn_exact
// Phase A: grid (chunks, row-group lanes).  A CTA keeps EX_Q query rows in shared
// memory and streams its pool chunk once; every warp scores one pool row against
// all EX_Q queries (fp64 accumulate) and keeps a per-(warp,query) sorted top-k.
// Phase B: one CTA per row merges chunks*k partials by (dist, idx).
constexpr int EX_Q = 8;
constexpr int EX_WARPS = 8;

struct ExEntry {
  double d;
  int64_t i;
};

__device__ __forceinline__ bool ex_less(double da, int64_t ia, double db, int64_t ib) {
  return da < db || (da == db && ia < ib);
}

__global__ void __launch_bounds__(EX_WARPS * 32) knn_exact_partial_kernel(
    const float* __restrict__ q, const double* __restrict__ qn, int64_t n_query, const float* __restrict__ p,
    const double* __restrict__ pn, int64_t n_pool, int dim, int k, const int64_t* __restrict__ row_list,
    const int* __restrict__ row_count_dev, int64_t row_count_host, int64_t slot_base, int64_t slot_cap,
    double* __restrict__ part_d, int64_t* __restrict__ part_i, int direct, int64_t index_offset,
    float* __restrict__ out_dist, double* __restrict__ out_dist64, int64_t* __restrict__ out_idx,
    const int64_t* __restrict__ mask_lo, const int64_t* __restrict__ mask_hi) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sq = reinterpret_cast<float*>(smem_raw);                       // [EX_Q][dim]
  double* ld = reinterpret_cast<double*>(sq + (size_t)EX_Q * dim);      // [EX_WARPS][EX_Q][k]
  int64_t* li = reinterpret_cast<int64_t*>(ld + EX_WARPS * EX_Q * k);   // same shape
  __shared__ int64_t rows_s[EX_Q];
  __shared__ double qn_s[EX_Q];
  __shared__ int64_t mlo_s[EX_Q], mhi_s[EX_Q];   // masked column range per query (distance := 1)

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_chunks = gridDim.x;
  const bool vec = (dim & 3) == 0 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(q) & 15) == 0);                      // the re-score kernel's own rule
  // Device-side row lists run in one of two regimes, chosen by the count the host cannot see:
  //   chunked (direct == 0): <= slot_cap rows, the pool is split over gridDim.x CTAs per row group
  //                          and per-chunk partial lists are merged by knn_exact_merge_kernel;
  //   direct  (direct == 1): > slot_cap rows, one CTA scans the whole pool for a row group and
  //                          writes the final lists (parallelism comes from the many groups).
  int64_t total = row_count_host;
  if (row_count_dev) {
    total = (int64_t)*row_count_dev;
    if ((total > slot_cap) != (direct != 0)) return;
  }
  const int64_t chunk_rows = ceil_div64(n_pool, n_chunks);
  const int64_t c0 = (int64_t)blockIdx.x * chunk_rows;
  const int64_t c1 = min(n_pool, c0 + chunk_rows);

  for (int64_t g = blockIdx.y; g * EX_Q < total; g += gridDim.y) {
    __syncthreads();
    if (threadIdx.x < EX_Q) {
      int64_t s = g * EX_Q + threadIdx.x;
      int64_t r = -1;
      if (s < total) r = row_list ? row_list[s] : (slot_base + s);
      rows_s[threadIdx.x] = r;
      qn_s[threadIdx.x] = (r >= 0) ? qn[r] : 1.0;
      mlo_s[threadIdx.x] = (r >= 0 && mask_lo) ? mask_lo[r] : 0;
      mhi_s[threadIdx.x] = (r >= 0 && mask_lo) ? mask_hi[r] : 0;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < EX_Q * dim; e += blockDim.x) {
      int qi = e / dim, c = e % dim;
      int64_t r = rows_s[qi];
      sq[e] = (r >= 0) ? __ldg(q + r * dim + c) : 0.f;
    }
    for (int e = threadIdx.x; e < EX_WARPS * EX_Q * k; e += blockDim.x) {
      ld[e] = INFINITY;
      li[e] = INT64_MAX;
    }
    __syncthreads();

    for (int64_t pr = c0 + warp; pr < c1; pr += EX_WARPS) {
      const float* prow = p + pr * dim;
      // The dot product follows knn_select.cu's rs_dot2 operation for operation (same element-to-lane
      // assignment, two fma chains per lane, same reduction tree): a (query, pool row) pair gets the
      // SAME fp64 bits whichever kernel scores it, so rows decided here and rows decided by the
      // re-score order exact duplicates identically (by index), also across pool shards.
      double acc[EX_Q];
      if (vec) {
        double a1[EX_Q];
#pragma unroll
        for (int qi = 0; qi < EX_Q; ++qi) {
          acc[qi] = 0.0;
          a1[qi] = 0.0;
        }
        const float4* p4 = reinterpret_cast<const float4*>(prow);
        for (int c = lane; c < dim / 4; c += 32) {
          const float4 v = __ldg(p4 + c);
#pragma unroll
          for (int qi = 0; qi < EX_Q; ++qi) {
            const float4 qv = *reinterpret_cast<const float4*>(sq + qi * dim + 4 * c);
            acc[qi] = fma((double)v.x, (double)qv.x, acc[qi]);
            a1[qi] = fma((double)v.y, (double)qv.y, a1[qi]);
            acc[qi] = fma((double)v.z, (double)qv.z, acc[qi]);
            a1[qi] = fma((double)v.w, (double)qv.w, a1[qi]);
          }
        }
#pragma unroll
        for (int qi = 0; qi < EX_Q; ++qi) acc[qi] += a1[qi];
      } else {
#pragma unroll
        for (int qi = 0; qi < EX_Q; ++qi) acc[qi] = 0.0;
        for (int c = lane; c < dim; c += 32) {
          const float pv = __ldg(prow + c);
#pragma unroll
          for (int qi = 0; qi < EX_Q; ++qi) acc[qi] = fma((double)pv, (double)sq[qi * dim + c], acc[qi]);
        }
      }
#pragma unroll
      for (int qi = 0; qi < EX_Q; ++qi) acc[qi] = warp_sum(acc[qi]);
      // lane qi owns query qi's list for this warp
      if (lane < EX_Q && rows_s[lane] >= 0) {
        double dot = 0.0;
#pragma unroll
        for (int qi = 0; qi < EX_Q; ++qi)
          if (qi == lane) dot = acc[qi];
        double d = 1.0 - dot / (qn_s[lane] * pn[pr]);
        if (pr >= mlo_s[lane] && pr < mhi_s[lane]) d = 1.0;
        double* ldq = ld + (warp * EX_Q + lane) * k;
        int64_t* liq = li + (warp * EX_Q + lane) * k;
        if (ex_less(d, pr, ldq[k - 1], liq[k - 1])) {
          int j = k - 1;
          while (j > 0 && ex_less(d, pr, ldq[j - 1], liq[j - 1])) {
            ldq[j] = ldq[j - 1];
            liq[j] = liq[j - 1];
            --j;
          }
          ldq[j] = d;
          liq[j] = pr;
        }
      }
    }
    __syncthreads();
    // merge the EX_WARPS lists of each query: warp w merges query w (EX_Q == EX_WARPS)
    {
      const int qi = warp;
      if (rows_s[qi] >= 0) {
        int64_t slot = g * EX_Q + qi;
        double* od = direct ? nullptr : part_d + ((slot * n_chunks) + blockIdx.x) * k;
        int64_t* oi = direct ? nullptr : part_i + ((slot * n_chunks) + blockIdx.x) * k;
        const int64_t orow = rows_s[qi];
        int head = 0;  // lanes 0..EX_WARPS-1: cursor into list of warp `lane`
        for (int o = 0; o < k; ++o) {
          double cd = INFINITY;
          int64_t ci = INT64_MAX;
          if (lane < EX_WARPS && head < k) {
            cd = ld[(lane * EX_Q + qi) * k + head];
            ci = li[(lane * EX_Q + qi) * k + head];
          }
          double bd = cd;
          int64_t bi = ci;
#pragma unroll
          for (int s = 16; s > 0; s >>= 1) {
            double od2 = __shfl_xor_sync(0xffffffffu, bd, s);
            int64_t oi2 = __shfl_xor_sync(0xffffffffu, bi, s);
            if (ex_less(od2, oi2, bd, bi)) {
              bd = od2;
              bi = oi2;
            }
          }
          if (lane < EX_WARPS && head < k && cd == bd && ci == bi && bi != INT64_MAX) ++head;
          if (lane == 0) {
            if (direct) {
              out_dist[orow * k + o] = (float)bd;
              if (out_dist64) out_dist64[orow * k + o] = bd;
              out_idx[orow * k + o] = (bi == INT64_MAX) ? -1 : bi + index_offset;
            } else {
              od[o] = bd;
              oi[o] = bi;
            }
          }
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256) knn_exact_merge_kernel(
    const double* __restrict__ part_d, const int64_t* __restrict__ part_i, int n_chunks, int k,
    const int64_t* __restrict__ row_list, const int* __restrict__ row_count_dev, int64_t row_count_host,
    int64_t slot_base, int64_t slot_cap, int64_t index_offset, float* __restrict__ out_dist,
    double* __restrict__ out_dist64, int64_t* __restrict__ out_idx) {
  int64_t total = row_count_host;
  if (row_count_dev) {
    total = (int64_t)*row_count_dev;
    if (total > slot_cap) return;   // the direct regime already wrote these rows
  }
  __shared__ double sd[256];
  __shared__ int64_t si[256];
  for (int64_t slot = blockIdx.x; slot < total; slot += gridDim.x) {
    const int64_t row = row_list ? row_list[slot] : (slot_base + slot);
    const double* pd = part_d + slot * n_chunks * k;
    const int64_t* pi = part_i + slot * n_chunks * k;
    const int n = n_chunks * k;
    double last_d = -INFINITY;
    int64_t last_i = -1;
    for (int o = 0; o < k; ++o) {
      double bd = INFINITY;
      int64_t bi = INT64_MAX;
      for (int e = threadIdx.x; e < n; e += blockDim.x) {
        double d = pd[e];
        int64_t i = pi[e];
        if (ex_less(last_d, last_i, d, i) && ex_less(d, i, bd, bi)) {
          bd = d;
          bi = i;
        }
      }
      sd[threadIdx.x] = bd;
      si[threadIdx.x] = bi;
      __syncthreads();
      for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s && ex_less(sd[threadIdx.x + s], si[threadIdx.x + s], sd[threadIdx.x], si[threadIdx.x])) {
          sd[threadIdx.x] = sd[threadIdx.x + s];
          si[threadIdx.x] = si[threadIdx.x + s];
        }
        __syncthreads();
      }
      last_d = sd[0];
      last_i = si[0];
      if (threadIdx.x == 0) {
        out_dist[row * k + o] = (float)last_d;
        if (out_dist64) out_dist64[row * k + o] = last_d;
        out_idx[row * k + o] = (last_i == INT64_MAX) ? -1 : last_i + index_offset;
      }
      __syncthreads();
    }
  }
}

int exact_chunks(int64_t n_pool) {
  int64_t c = ceil_div64(n_pool, 2048);
  if (c < 1) c = 1;
  if (c > 64) c = 64;
  return (int)c;
}

size_t exact_partial_bytes(int64_t slots, int64_t n_pool, int k) {
  return (size_t)slots * exact_chunks(n_pool) * k * (sizeof(double) + sizeof(int64_t));
}

// Runs the exact kNN for `slots` rows: either the device-side list (row_list,
// row_count_dev; at most slot_cap rows) or the contiguous rows [slot_base, slot_base+row_count_host).
int launch_knn_exact_rows(const float* q, const double* qn, int64_t n_query, const float* p, const double* pn,
                          int64_t n_pool, int dim, int k, const int64_t* row_list, const int* row_count_dev,
                          int64_t row_count_host, int64_t slot_base, int64_t slot_cap, int64_t index_offset,
                          float* out_dist, double* out_dist64, int64_t* out_idx, void* partial,
                          const int64_t* mask_lo, const int64_t* mask_hi, cudaStream_t stream) {
  const int n_chunks = exact_chunks(n_pool);
  double* part_d = reinterpret_cast<double*>(partial);
  int64_t* part_i = reinterpret_cast<int64_t*>(part_d + (size_t)slot_cap * n_chunks * k);
  const int64_t max_groups = ceil_div64(row_count_dev ? slot_cap : row_count_host, EX_Q);
  int gy = (int)(max_groups < 1 ? 1 : max_groups);
  const int target = 148 * 4;
  if ((int64_t)gy * n_chunks > target) gy = (target + n_chunks - 1) / n_chunks;
  if (gy < 1) gy = 1;
  size_t smem = (size_t)EX_Q * dim * sizeof(float) + (size_t)EX_WARPS * EX_Q * k * (sizeof(double) + sizeof(int64_t));
  KNN_CHECK_ARG(smem <= 200 * 1024, -2, "knn_exact: dim %d too large for shared memory", dim);
  static PerDevice attr;
  KNN_SMEM_ATTR(attr, knn_exact_partial_kernel, 200 * 1024);
  dim3 grid(n_chunks, gy);
  knn_exact_partial_kernel<<<grid, EX_WARPS * 32, smem, stream>>>(q, qn, n_query, p, pn, n_pool, dim, k, row_list,
                                                                   row_count_dev, row_count_host, slot_base,
                                                                   slot_cap, part_d, part_i, 0, index_offset, out_dist,
                                                                   out_dist64, out_idx, mask_lo, mask_hi);
  KNN_LAUNCH_CHECK();
  if (row_count_dev) {
    // many-rows regime (exits immediately unless the device-side count exceeds slot_cap)
    knn_exact_partial_kernel<<<dim3(1, 148 * 2), EX_WARPS * 32, smem, stream>>>(
        q, qn, n_query, p, pn, n_pool, dim, k, row_list, row_count_dev, row_count_host, slot_base, slot_cap, part_d,
        part_i, 1, index_offset, out_dist, out_dist64, out_idx, mask_lo, mask_hi);
    KNN_LAUNCH_CHECK();
  }
  int64_t mg = row_count_dev ? slot_cap : row_count_host;
  if (mg > 148 * 8) mg = 148 * 8;
  if (mg < 1) mg = 1;
  knn_exact_merge_kernel<<<(unsigned)mg, 256, 0, stream>>>(part_d, part_i, n_chunks, k, row_list, row_count_dev,
                                                           row_count_host, slot_base, slot_cap, index_offset,
                                                           out_dist, out_dist64, out_idx);
  KNN_LAUNCH_CHECK();
  return 0;
}

}  // namespace knnsvc
