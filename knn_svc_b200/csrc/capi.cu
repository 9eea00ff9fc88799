// extern "C" entry points declared in include/knnsvc_b200.h.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "../../include/knnsvc_b200.h"
#include "common.cuh"
#include "kernels.cuh"

namespace knnsvc {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// Tuning options: process-wide diagnostic switches (knnsvc_set_option).  Each is one relaxed atomic
// read at launch time, so setting one from another thread is well defined; none of them changes
// results, only which kernel shape / schedule produces them.
static std::atomic<int> g_opt_bf16{0};
int opt_bf16() { return g_opt_bf16.load(std::memory_order_relaxed); }
static std::atomic<int> g_opt_filter_flags{1};
int opt_filter_flags() { return g_opt_filter_flags.load(std::memory_order_relaxed); }
static std::atomic<int> g_opt_wf_cluster{1};
int opt_weight_fit_cluster() { return g_opt_wf_cluster.load(std::memory_order_relaxed); }
static std::atomic<int> g_opt_block_tiles{0};
int opt_block_tiles() { return g_opt_block_tiles.load(std::memory_order_relaxed); }
static std::atomic<int> g_opt_spin_ns{40};  // measured: ~3% faster than a pure spin under the power cap
int opt_spin_ns() { return g_opt_spin_ns.load(std::memory_order_relaxed); }
static std::atomic<int> g_opt_epi_sleep_ns{0};
int opt_epi_sleep_ns() { return g_opt_epi_sleep_ns.load(std::memory_order_relaxed); }

static std::atomic<int> g_opt_refine_min{0};
int opt_refine_min() { return g_opt_refine_min.load(std::memory_order_relaxed); }
static std::atomic<int> g_opt_query_group{0};
int opt_query_group() { return g_opt_query_group.load(std::memory_order_relaxed); }
static std::atomic<int> g_opt_log_cap{0};
int opt_log_cap() { return g_opt_log_cap.load(std::memory_order_relaxed); }

static std::atomic<int> g_opt_concat_staged{1};
int opt_concat_staged() { return g_opt_concat_staged.load(std::memory_order_relaxed); }

static std::atomic<int> g_opt_concat_cluster{1};
int opt_concat_cluster() { return g_opt_concat_cluster.load(std::memory_order_relaxed); }

static std::atomic<int> g_opt_concat_f0_table{1};
int opt_concat_f0_table() { return g_opt_concat_f0_table.load(std::memory_order_relaxed); }

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// Optional device-side timing of the dominant kernel (the tcgen05 filter): when
// enabled, knn_search brackets that one launch with CUDA events on its stream.
// The event table is guarded by a mutex (searches may come from several host threads).
constexpr int kTimingSlots = 256;
static std::mutex g_timing_mu;
static bool g_timing = false;
static cudaEvent_t g_ev[kTimingSlots][2];
static bool g_ev_made = false;
static int g_ev_n = 0;

namespace {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct KnnWorkspace {
  float* log_val;
  int* log_idx;
  int* log_cnt;
  float* seg_top;
  float* seg_kth;
  float* ref_val;   // refined (fp32) similarity of every log entry, -inf below the first threshold
  float* row_thr;   // first threshold of every row (+inf: the row overflowed its log)
  int* row_mode;    // decision route of every row (1: fp32 refine first, 0: straight to fp64)
  int* blk_off;     // [slots][rf_nblk + 1] start of every refine block inside the (sorted) log
  int* seg_flag;
  size_t seg_flag_bytes;
  int64_t* flag_list;
  int* counters;  // [0] flag count, [1..8] stats
  void* exact_partial;
  size_t total;
};

KnnWorkspace carve(void* base, int64_t n_query, int64_t n_pool, int k, const FilterPlan& pl) {
  // (log_val .. seg_top come first and do not depend on the feature dimension: knnsvc_knn_workspace_layout)
  KnnWorkspace w;
  size_t off = 0;
  unsigned char* b = reinterpret_cast<unsigned char*>(base);
  const size_t slots = (size_t)n_query * pl.n_seg;
  auto take = [&](size_t bytes) {
    void* p = b ? b + off : nullptr;
    off += align_up(bytes, 256);
    return p;
  };
  w.log_val = reinterpret_cast<float*>(take(slots * pl.cap * sizeof(float)));
  w.log_idx = reinterpret_cast<int*>(take(slots * pl.cap * sizeof(int)));
  w.log_cnt = reinterpret_cast<int*>(take(slots * sizeof(int)));
  w.seg_top = reinterpret_cast<float*>(take(slots * k * sizeof(float)));
  w.seg_kth = reinterpret_cast<float*>(take(slots * sizeof(float)));
  w.ref_val = reinterpret_cast<float*>(take(slots * pl.cap * sizeof(float)));
  w.row_thr = reinterpret_cast<float*>(take((size_t)n_query * sizeof(float)));
  w.row_mode = reinterpret_cast<int*>(take((size_t)n_query * sizeof(int)));
  w.blk_off = reinterpret_cast<int*>(take(slots * (size_t)(pl.rf_nblk + 1) * sizeof(int)));
  w.seg_flag_bytes = filter_flag_count(pl) * sizeof(int);
  w.seg_flag = reinterpret_cast<int*>(take(w.seg_flag_bytes));
  w.flag_list = reinterpret_cast<int64_t*>(take((size_t)n_query * sizeof(int64_t)));
  w.counters = reinterpret_cast<int*>(take(16 * sizeof(int)));
  w.exact_partial = take(exact_partial_bytes(kFlagCap, n_pool, k));
  w.total = off;
  return w;
}

FilterPlan full_plan(int64_t n_query, int64_t n_pool, int dim_pad, int k) {
  FilterPlan pl = plan_filter(n_query, n_pool, k);
  plan_refine(n_query, n_pool, dim_pad, &pl.rf_rows, &pl.rf_nblk);
  return pl;
}

__global__ void write_plan_stats(int* stats, const int* counters, int n_seg, int n_units, int grid, int cap) {
  stats[0] = counters[1];
  stats[1] = counters[2];
  stats[2] = counters[3];
  stats[3] = n_seg;
  stats[4] = n_units;
  stats[5] = grid;
  stats[6] = cap;
  stats[7] = counters[7];   // candidates inside the refined window (scored in fp64)
}

}  // namespace
}  // namespace knnsvc

using namespace knnsvc;

// One stream-ordered memory pool per device for the library's few internal scratch buffers.
static int scratch_pool(cudaMemPool_t* out) {
  static cudaMemPool_t pools[64] = {};
  static std::mutex mu;
  int dev = 0;
  KNN_CUDA(cudaGetDevice(&dev));
  KNN_CHECK_ARG(dev >= 0 && dev < 64, -1, "device index %d out of range", dev);
  std::lock_guard<std::mutex> lock(mu);
  if (!pools[dev]) {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    cudaMemPool_t pool_h = nullptr;
    KNN_CUDA(cudaMemPoolCreate(&pool_h, &props));
    unsigned long long keep = ~0ull;
    KNN_CUDA(cudaMemPoolSetAttribute(pool_h, cudaMemPoolAttrReleaseThreshold, &keep));
    pools[dev] = pool_h;
  }
  *out = pools[dev];
  return 0;
}

extern "C" {

const char* knnsvc_last_error(void) { return g_err; }
int knnsvc_version(void) { return 100; }

int knnsvc_prepare_rows(const float* x, int64_t rows, int dim, int64_t ld, void* half_out, int dim_pad, double* norms,
                        int* bad_rows, float* max_err, void* stream) {
  KNN_CHECK_ARG(rows >= 0 && dim >= 1 && ld >= dim && dim_pad >= dim, -1, "prepare_rows: bad shape");
  if (rows == 0) return 0;   // an empty row set has no storage to point at
  KNN_CHECK_ARG(x && half_out && norms && bad_rows, -1, "prepare_rows: null pointer");
  return launch_prepare_rows(x, rows, dim, ld, half_out, dim_pad, norms, bad_rows, max_err, (cudaStream_t)stream);
}

int knnsvc_cosine_dist(const float* q, int64_t n_query, const float* p, int64_t n_pool, int dim, float* out,
                       void* stream) {
  KNN_CHECK_ARG(n_query >= 0 && n_pool >= 0 && dim >= 1, -1, "cosine_dist: bad shape");
  return launch_cosine_dist(q, n_query, p, n_pool, dim, out, (cudaStream_t)stream);
}

size_t knnsvc_knn_workspace_bytes(int64_t n_query, int64_t n_pool, int dim_pad, int k) {
  if (n_query <= 0 || n_pool <= 0 || k < 1 || k > kMaxK || dim_pad < 1) return 0;
  FilterPlan pl = full_plan(n_query, n_pool, dim_pad, k);
  return carve(nullptr, n_query, n_pool, k, pl).total;
}

int knnsvc_knn_plan(int64_t n_query, int64_t n_pool, int k, int* plan_host) {
  KNN_CHECK_ARG(plan_host != nullptr, -1, "knn_plan: null output");
  KNN_CHECK_ARG(n_query >= 1 && n_pool >= 1 && k >= 1 && k <= kMaxK, -1, "knn_plan: bad shape");
  const FilterPlan pl = plan_filter(n_query, n_pool, k);
  const int v[8] = {pl.ctas, pl.n_qtiles, pl.n_ptiles, pl.n_seg, pl.n_blk, pl.n_units, pl.grid, pl.cap};
  for (int i = 0; i < 8; ++i) plan_host[i] = v[i];
  return 0;
}

int knnsvc_knn_search(const float* q, const void* qh, const double* qn, int64_t n_query, const float* p,
                      const void* ph, const double* pn, int64_t n_pool, int dim, int dim_pad, int k,
                      int64_t index_offset, const float* q_err, const float* p_err, float* out_dist,
                      int64_t* out_idx, void* workspace, size_t workspace_bytes, int* stats, void* stream_) {
  return knnsvc_knn_search_masked(q, qh, qn, n_query, p, ph, pn, n_pool, dim, dim_pad, k, index_offset, q_err, p_err,
                                  nullptr, nullptr, out_dist, out_idx, workspace, workspace_bytes, stats, stream_);
}

int knnsvc_knn_search_masked(const float* q, const void* qh, const double* qn, int64_t n_query, const float* p,
                             const void* ph, const double* pn, int64_t n_pool, int dim, int dim_pad, int k,
                             int64_t index_offset, const float* q_err, const float* p_err,
                             const int64_t* mask_lo, const int64_t* mask_hi, float* out_dist, int64_t* out_idx,
                             void* workspace, size_t workspace_bytes, int* stats, void* stream_) {
  return knnsvc_knn_search_full(q, qh, qn, n_query, p, ph, pn, n_pool, dim, dim_pad, k, index_offset, q_err, p_err,
                                mask_lo, mask_hi, out_dist, nullptr, out_idx, workspace, workspace_bytes, stats,
                                stream_);
}

int knnsvc_knn_search_full(const float* q, const void* qh, const double* qn, int64_t n_query, const float* p,
                           const void* ph, const double* pn, int64_t n_pool, int dim, int dim_pad, int k,
                           int64_t index_offset, const float* q_err, const float* p_err, const int64_t* mask_lo,
                           const int64_t* mask_hi, float* out_dist, double* out_dist64, int64_t* out_idx,
                           void* workspace, size_t workspace_bytes, int* stats, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  KNN_CHECK_ARG((mask_lo == nullptr) == (mask_hi == nullptr), -1,
                "knn_search: mask_lo and mask_hi must be given together");
  KNN_CHECK_ARG(n_query >= 0 && n_pool >= 1 && dim >= 1 && dim_pad >= dim, -1, "knn_search: bad shape");
  KNN_CHECK_ARG(k >= 1 && k <= kMaxK, -1, "knn_search: k=%d outside [1,%d]", k, kMaxK);
  KNN_CHECK_ARG(k <= n_pool, -1, "knn_search: k=%d exceeds the pool size %lld", k, (long long)n_pool);
  if (n_query == 0) return 0;
  KNN_CHECK_ARG(q && qh && qn && p && ph && pn && out_dist && out_idx && workspace, -1, "knn_search: null pointer");
  KNN_CHECK_ARG(!opt_bf16() || (q_err && p_err), -1,
                "knn_search: bf16 operands need the measured row errors (the default window assumes fp16)");
  FilterPlan pl = full_plan(n_query, n_pool, dim_pad, k);
  KnnWorkspace w = carve(workspace, n_query, n_pool, k, pl);
  KNN_CHECK_ARG(workspace_bytes >= w.total, -2, "knn_search: workspace %zu < required %zu", workspace_bytes, w.total);
  KNN_CUDA(cudaMemsetAsync(w.counters, 0, 16 * sizeof(int), stream));
  KNN_CUDA(cudaMemsetAsync(w.seg_flag, 0, w.seg_flag_bytes, stream));
  int ev_slot = -1;
  {
    std::lock_guard<std::mutex> lock(g_timing_mu);
    if (g_timing && g_ev_n < kTimingSlots) ev_slot = g_ev_n++;
  }
  const bool timed = ev_slot >= 0;
  if (timed) KNN_CUDA(cudaEventRecord(g_ev[ev_slot][0], stream));
  int rc = launch_knn_filter(qh, n_query, ph, n_pool, dim_pad, k, pl, w.log_val, w.log_idx, w.log_cnt, w.seg_top,
                             w.seg_kth, w.seg_flag, mask_lo, mask_hi, q_err, p_err, w.counters + 12, stream);
  if (rc) return rc;
  if (timed) KNN_CUDA(cudaEventRecord(g_ev[ev_slot][1], stream));
  rc = launch_knn_rescore(q, qn, n_query, p, pn, n_pool, dim, k, pl, w.log_val, w.log_idx, w.log_cnt, w.seg_top,
                          w.ref_val, w.row_thr, w.row_mode, w.blk_off, index_offset, out_dist, out_dist64, out_idx, w.flag_list,
                          w.counters, w.counters + 1, mask_lo, mask_hi, q_err, p_err, stream);
  if (rc) return rc;
  // rows the error window could not decide: exact brute force, count known only on the device
  rc = launch_knn_exact_rows(q, qn, n_query, p, pn, n_pool, dim, k, w.flag_list, w.counters, 0, 0, kFlagCap,
                             index_offset, out_dist, out_dist64, out_idx, w.exact_partial, mask_lo, mask_hi,
                             stream);
  if (rc) return rc;
  if (stats) {
    write_plan_stats<<<1, 1, 0, stream>>>(stats, w.counters, pl.n_seg, pl.n_units, pl.grid, pl.cap);
    KNN_LAUNCH_CHECK();
  }
  return 0;
}

long long knnsvc_launch_count(void) { return g_launches.load(); }

int knnsvc_set_option(const char* name, int value) {
  KNN_CHECK_ARG(name != nullptr, -1, "set_option: null name");
  if (strcmp(name, "cta_group") == 0) {
    KNN_CHECK_ARG(value == 1, -1, "set_option: the cta_group::2 variant was removed (measured slower); only 1 is accepted");
    return 0;
  }
  if (strcmp(name, "spin_sleep_ns") == 0) {
    KNN_CHECK_ARG(value >= 0 && value <= 60000, -1, "set_option: spin_sleep_ns out of range");
    g_opt_spin_ns.store(value, std::memory_order_relaxed);
    return 0;
  }
  if (strcmp(name, "epi_sleep_ns") == 0) {
    KNN_CHECK_ARG(value >= 0 && value <= 60000, -1, "set_option: epi_sleep_ns out of range");
    g_opt_epi_sleep_ns.store(value, std::memory_order_relaxed);
    return 0;
  }
  if (strcmp(name, "concat_staged") == 0) {
    g_opt_concat_staged.store(value != 0, std::memory_order_relaxed);
    return 0;
  }
  if (strcmp(name, "concat_cluster") == 0) {
    g_opt_concat_cluster.store(value != 0, std::memory_order_relaxed);
    return 0;
  }
  if (strcmp(name, "concat_f0_table") == 0) {
    g_opt_concat_f0_table.store(value != 0, std::memory_order_relaxed);
    return 0;
  }
  if (strcmp(name, "filter_flags") == 0) {
    KNN_CHECK_ARG(value >= 0 && value <= 7, -1, "set_option: filter_flags out of range");
    g_opt_filter_flags.store(value, std::memory_order_relaxed);
    return 0;
  }
  if (strcmp(name, "weight_fit_cluster") == 0) {
    g_opt_wf_cluster.store(value != 0, std::memory_order_relaxed);
    return 0;
  }
  if (strcmp(name, "block_tiles") == 0) {
    KNN_CHECK_ARG(value >= 0 && value <= (1 << 20), -1, "set_option: block_tiles out of range");
    g_opt_block_tiles.store(value, std::memory_order_relaxed);
    return 0;
  }
  if (strcmp(name, "refine_min_candidates") == 0) {
    KNN_CHECK_ARG(value >= 0 && value <= (1 << 20), -1, "set_option: refine_min_candidates out of range");
    g_opt_refine_min.store(value, std::memory_order_relaxed);
    return 0;
  }
  if (strcmp(name, "query_group") == 0) {
    KNN_CHECK_ARG(value >= 0 && value <= (1 << 24), -1, "set_option: query_group out of range");
    g_opt_query_group.store(value, std::memory_order_relaxed);
    return 0;
  }
  if (strcmp(name, "log_cap") == 0) {
    KNN_CHECK_ARG(value == 0 || (value >= 64 && value <= 65536 && value % 8 == 0), -1,
                  "set_option: log_cap must be 0 (default) or a multiple of 8 in [64, 65536]");
    g_opt_log_cap.store(value, std::memory_order_relaxed);
    return 0;
  }
  if (strcmp(name, "bf16_operands") == 0) {
    g_opt_bf16.store(value != 0, std::memory_order_relaxed);
    return 0;
  }
  KNN_CHECK_ARG(false, -1, "set_option: unknown option '%s'", name);
}

int knnsvc_filter_timing(int enable) {
  std::lock_guard<std::mutex> lock(g_timing_mu);
  if (enable && !g_ev_made) {
    for (int i = 0; i < kTimingSlots; ++i)
      for (int j = 0; j < 2; ++j) KNN_CUDA(cudaEventCreate(&g_ev[i][j]));
    g_ev_made = true;
  }
  g_timing = enable != 0;
  g_ev_n = 0;
  return 0;
}

int knnsvc_filter_timing_collect(float* ms_host, int max_n) {
  std::lock_guard<std::mutex> lock(g_timing_mu);
  int n = g_ev_n < max_n ? g_ev_n : max_n;
  for (int i = 0; i < n; ++i) {
    if (cudaEventSynchronize(g_ev[i][1]) != cudaSuccess) return -1;
    if (cudaEventElapsedTime(ms_host + i, g_ev[i][0], g_ev[i][1]) != cudaSuccess) return -1;
  }
  g_ev_n = 0;
  return n;
}

size_t knnsvc_knn_exact_workspace_bytes(int64_t n_query, int64_t n_pool, int k) {
  int64_t slots = n_query < kFlagCap ? n_query : kFlagCap;
  if (slots < 1) slots = 1;
  return exact_partial_bytes(slots, n_pool, k) + 256;
}

int knnsvc_knn_exact(const float* q, const double* qn, int64_t n_query, const float* p, const double* pn,
                     int64_t n_pool, int dim, int k, int64_t index_offset, float* out_dist, int64_t* out_idx,
                     void* workspace, size_t workspace_bytes, void* stream) {
  KNN_CHECK_ARG(n_query >= 0 && n_pool >= 1 && dim >= 1, -1, "knn_exact: bad shape");
  KNN_CHECK_ARG(k >= 1 && k <= kMaxK && k <= n_pool, -1, "knn_exact: bad k=%d", k);
  KNN_CHECK_ARG(workspace_bytes >= knnsvc_knn_exact_workspace_bytes(n_query, n_pool, k), -2,
                "knn_exact: workspace too small");
  const int64_t cap = n_query < kFlagCap ? n_query : kFlagCap;
  for (int64_t base = 0; base < n_query; base += cap) {
    const int64_t n = (n_query - base) < cap ? (n_query - base) : cap;
    int rc = launch_knn_exact_rows(q, qn, n_query, p, pn, n_pool, dim, k, nullptr, nullptr, n, base, cap,
                                   index_offset, out_dist, nullptr, out_idx, workspace, nullptr, nullptr,
                                   (cudaStream_t)stream);
    if (rc) return rc;
  }
  return 0;
}

int knnsvc_merge_topk(const float* gathered_dist, const int64_t* gathered_idx, int n_shards, int64_t n_query, int k,
                      float* out_dist, int64_t* out_idx, void* stream) {
  KNN_CHECK_ARG(gathered_dist && gathered_idx && out_dist && out_idx, -1, "merge_topk: null pointer");
  return launch_merge_topk(gathered_dist, gathered_idx, n_shards, n_query, k, out_dist, out_idx, (cudaStream_t)stream);
}

int knnsvc_merge_topk64(const double* gathered_dist, const int64_t* gathered_idx, int n_shards, int64_t n_query,
                        int k, float* out_dist, double* out_dist64, int64_t* out_idx, void* stream) {
  KNN_CHECK_ARG(gathered_dist && gathered_idx && out_dist && out_idx, -1, "merge_topk64: null pointer");
  return launch_merge_topk64(gathered_dist, gathered_idx, n_shards, n_query, k, out_dist, out_dist64, out_idx,
                             (cudaStream_t)stream);
}

int knnsvc_knn_workspace_layout(int64_t n_query, int64_t n_pool, int dim_pad, int k, int64_t* layout_host) {
  KNN_CHECK_ARG(layout_host != nullptr, -1, "knn_workspace_layout: null output");
  KNN_CHECK_ARG(n_query >= 1 && n_pool >= 1 && k >= 1 && k <= kMaxK && dim_pad >= 1, -1,
                "knn_workspace_layout: bad shape");
  const FilterPlan pl = full_plan(n_query, n_pool, dim_pad, k);
  unsigned char* base = reinterpret_cast<unsigned char*>(uintptr_t(1) << 20);   // any non-null base: offsets only
  const KnnWorkspace w = carve(base, n_query, n_pool, k, pl);
  layout_host[0] = reinterpret_cast<unsigned char*>(w.log_val) - base;
  layout_host[1] = reinterpret_cast<unsigned char*>(w.log_idx) - base;
  layout_host[2] = reinterpret_cast<unsigned char*>(w.log_cnt) - base;
  layout_host[3] = reinterpret_cast<unsigned char*>(w.seg_top) - base;
  layout_host[4] = pl.n_seg;
  layout_host[5] = pl.cap;
  layout_host[6] = (int64_t)w.total;
  layout_host[7] = reinterpret_cast<unsigned char*>(w.ref_val) - base;
  return 0;
}

// ---- peer memory (one process per GPU): CUDA IPC handles of a shard's row storage
int knnsvc_ipc_export(const void* ptr, void* handle_host, int64_t* offset_host) {
  KNN_CHECK_ARG(ptr && handle_host && offset_host, -1, "ipc_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* base = nullptr;
  size_t size = 0;
  // the handle names the whole ALLOCATION the pointer lies in; the caller's tensor may start inside it
  typedef CUresult (*RangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  KNN_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres));   // no link-time libcuda
  KNN_CHECK_ARG(fn != nullptr && qres == cudaDriverEntryPointSuccess, -10, "cuMemGetAddressRange entry point not available");
  CUresult r = reinterpret_cast<RangeFn>(fn)(reinterpret_cast<CUdeviceptr*>(&base), &size,
                                             reinterpret_cast<CUdeviceptr>(ptr));
  KNN_CHECK_ARG(r == CUDA_SUCCESS, -12, "ipc_export: cuMemGetAddressRange failed with CUresult %d", (int)r);
  cudaIpcMemHandle_t h;
  KNN_CUDA(cudaIpcGetMemHandle(&h, base));
  memcpy(handle_host, &h, sizeof(h));
  *offset_host = (int64_t)(reinterpret_cast<const unsigned char*>(ptr) - reinterpret_cast<const unsigned char*>(base));
  return 0;
}

int knnsvc_ipc_open(const void* handle_host, void** base_out) {
  KNN_CHECK_ARG(handle_host && base_out, -1, "ipc_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_host, sizeof(h));
  KNN_CUDA(cudaIpcOpenMemHandle(base_out, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

int knnsvc_ipc_close(void* base) {
  if (!base) return 0;
  KNN_CUDA(cudaIpcCloseMemHandle(base));
  return 0;
}

static int make_table(const void* const* shard_rows_host, const int64_t* shard_lo_host, int n_shards, RowTable* tab,
                      const char* what) {
  KNN_CHECK_ARG(shard_rows_host && shard_lo_host, -1, "%s: null shard table", what);
  KNN_CHECK_ARG(n_shards >= 1 && n_shards <= kMaxShards, -1, "%s: %d shards outside [1,%d]", what, n_shards, kMaxShards);
  tab->n = n_shards;
  for (int s = 0; s < n_shards; ++s) {
    KNN_CHECK_ARG(shard_rows_host[s] != nullptr && shard_lo_host[s + 1] >= shard_lo_host[s], -1, "%s: bad shard %d", what,
                  s);
    tab->base[s] = reinterpret_cast<const float*>(shard_rows_host[s]);
    tab->lo[s] = shard_lo_host[s];
  }
  tab->lo[n_shards] = shard_lo_host[n_shards];
  KNN_CHECK_ARG(shard_lo_host[0] == 0 && tab->lo[n_shards] >= 1, -1, "%s: the shards must cover rows [0, n) with n >= 1", what);
  return 0;
}

int knnsvc_gather_mix_sharded(const void* const* shard_rows_host, const int64_t* shard_lo_host, int n_shards, int dim,
                              const int64_t* idx, const float* weights, int64_t n_query, int k, float* out,
                              void* stream) {
  KNN_CHECK_ARG(idx && out && k >= 1, -1, "gather_mix_sharded: bad arguments");
  RowTable tab;
  int rc = make_table(shard_rows_host, shard_lo_host, n_shards, &tab, "gather_mix_sharded");
  if (rc) return rc;
  return launch_gather_mix_sharded(tab, dim, idx, weights, n_query, k, out, (cudaStream_t)stream);
}

int knnsvc_gather_mix(const float* pool, int64_t n_pool, int dim, const int64_t* idx, const float* weights,
                      int64_t n_query, int k, float* out, void* stream) {
  KNN_CHECK_ARG(pool && idx && out && n_pool >= 1 && k >= 1, -1, "gather_mix: bad arguments");
  return launch_gather_mix(pool, n_pool, dim, idx, weights, n_query, k, out, (cudaStream_t)stream);
}

int knnsvc_f0_rerank(const float* expected_f0, const float* pool_f0, const int64_t* idx, int64_t n_query, int k,
                     int64_t* out_idx, void* stream) {
  KNN_CHECK_ARG(expected_f0 && pool_f0 && idx && out_idx, -1, "f0_rerank: null pointer");
  return launch_f0_rerank(expected_f0, pool_f0, idx, n_query, k, out_idx, (cudaStream_t)stream);
}

static int concat_cost_on_table(const int64_t* idx, const float* src, const RowTable& pool, int dim,
                                const float* shifted_src_f0, const float* pool_f0, float concat_weight,
                                const int64_t* utt_offsets_host, int n_utt, int64_t* out_idx, void* stream_);

int knnsvc_concat_cost_reselect(const int64_t* idx, const float* src, const float* pool, int64_t n_pool, int dim,
                                const float* shifted_src_f0, const float* pool_f0, float concat_weight,
                                const int64_t* utt_offsets_host, int n_utt, int64_t* out_idx, void* stream_) {
  KNN_CHECK_ARG(pool && n_pool >= 1, -1, "concat_cost: bad arguments");
  return concat_cost_on_table(idx, src, single_table(pool, n_pool), dim, shifted_src_f0, pool_f0, concat_weight,
                              utt_offsets_host, n_utt, out_idx, stream_);
}

int knnsvc_concat_cost_reselect_sharded(const int64_t* idx, const float* src, const void* const* shard_rows_host,
                                        const int64_t* shard_lo_host, int n_shards, int dim,
                                        const float* shifted_src_f0, const float* pool_f0, float concat_weight,
                                        const int64_t* utt_offsets_host, int n_utt, int64_t* out_idx, void* stream_) {
  RowTable tab;
  int rc = make_table(shard_rows_host, shard_lo_host, n_shards, &tab, "concat_cost_sharded");
  if (rc) return rc;
  return concat_cost_on_table(idx, src, tab, dim, shifted_src_f0, pool_f0, concat_weight, utt_offsets_host, n_utt,
                              out_idx, stream_);
}

static int concat_cost_on_table(const int64_t* idx, const float* src, const RowTable& pool, int dim,
                                const float* shifted_src_f0, const float* pool_f0, float concat_weight,
                                const int64_t* utt_offsets_host, int n_utt, int64_t* out_idx, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  KNN_CHECK_ARG(idx && src && out_idx && utt_offsets_host && n_utt >= 0, -1, "concat_cost: bad arguments");
  KNN_CHECK_ARG((shifted_src_f0 == nullptr) == (pool_f0 == nullptr), -1,
                "concat_cost: shifted_src_f0 and pool_f0 must be given together");
  if (n_utt == 0) return 0;
  const int64_t n_frames = utt_offsets_host[n_utt];
  KNN_CHECK_ARG(utt_offsets_host[0] == 0 && n_frames >= 0, -1, "concat_cost: utterance offsets must start at 0");
  // stream-ordered scratch: utterance offsets + per-frame baseline and |src|^2 (fp64).  It comes from
  // a private pool that KEEPS its memory across synchronisations: the default pool's release
  // threshold is 0, so a caller that synchronises after every utterance paid a fresh physical
  // allocation (~3 ms) on every call.
  unsigned char* d_ws = nullptr;
  const size_t off_bytes = ((size_t)(n_utt + 1) * sizeof(int64_t) + 255) / 256 * 256;
  cudaMemPool_t pool_h = nullptr;
  int rc_pool = scratch_pool(&pool_h);
  if (rc_pool) return rc_pool;
  // (+ for f0 runs of a few utterances against a pool of moderate size: the table of log2 f0 of the cluster kernel)
  const int64_t n_pool_rows = pool.lo[pool.n];
  const bool lf0_table = shifted_src_f0 != nullptr && n_utt <= 32 && n_pool_rows <= ((int64_t)1 << 22) &&
                         opt_concat_f0_table();
  const size_t frame_bytes = (size_t)2 * (n_frames + 1) * sizeof(double);
  KNN_CUDA(cudaMallocFromPoolAsync(reinterpret_cast<void**>(&d_ws),
                                   off_bytes + frame_bytes + (lf0_table ? (size_t)n_pool_rows * sizeof(double) : 0),
                                   pool_h, stream));
  int64_t* d_off = reinterpret_cast<int64_t*>(d_ws);
  KNN_CUDA(cudaMemcpyAsync(d_off, utt_offsets_host, (size_t)(n_utt + 1) * sizeof(int64_t), cudaMemcpyHostToDevice,
                           stream));
  int rc = launch_concat_cost(idx, src, pool, dim, shifted_src_f0, pool_f0, concat_weight, d_off, n_utt,
                              n_frames, reinterpret_cast<double*>(d_ws + off_bytes),
                              lf0_table ? reinterpret_cast<double*>(d_ws + off_bytes + frame_bytes) : nullptr, out_idx,
                              stream);
  cudaFreeAsync(d_ws, stream);
  return rc;
}

size_t knnsvc_weight_fit_workspace_bytes(int64_t n_query, int k) { return weight_fit_workspace_bytes(n_query, k, 1); }

int knnsvc_weight_fit(const int64_t* idx, const float* synth, int64_t n_pool, int dim, int64_t n_query, int k,
                      double loss_scale, int max_iters, float* out_weights, double* info, void* workspace,
                      size_t workspace_bytes, void* stream) {
  KNN_CHECK_ARG(idx && synth && out_weights && workspace && n_query >= 0, -1, "weight_fit: bad arguments");
  KNN_CHECK_ARG(workspace_bytes >= weight_fit_workspace_bytes(n_query, k, 1), -2, "weight_fit: workspace too small");
  const int64_t offs[2] = {0, n_query};
  return launch_weight_fit(idx, single_table(synth, n_pool), dim, offs, 1, k, loss_scale, max_iters, nullptr, out_weights,
                           info, workspace, (cudaStream_t)stream);
}

size_t knnsvc_weight_fit_batched_workspace_bytes(int64_t n_frames, int k, int n_utt) {
  return weight_fit_workspace_bytes(n_frames, k, n_utt);
}

int knnsvc_weight_fit_batched(const int64_t* idx, const float* synth, int64_t n_pool, int dim,
                              const int64_t* utt_offsets_host, int n_utt, int k, double loss_scale, int max_iters,
                              float* out_weights, double* info, void* workspace, size_t workspace_bytes,
                              void* stream) {
  return knnsvc_weight_fit_amp(idx, synth, n_pool, dim, utt_offsets_host, n_utt, k, loss_scale, max_iters, nullptr,
                               out_weights, info, workspace, workspace_bytes, stream);
}

int knnsvc_weight_fit_amp(const int64_t* idx, const float* synth, int64_t n_pool, int dim,
                          const int64_t* utt_offsets_host, int n_utt, int k, double loss_scale, int max_iters,
                          const float* amp_ratio, float* out_weights, double* info, void* workspace,
                          size_t workspace_bytes, void* stream) {
  KNN_CHECK_ARG(idx && synth && out_weights && workspace && utt_offsets_host && n_utt >= 0, -1,
                "weight_fit: bad arguments");
  if (n_utt == 0) return 0;
  KNN_CHECK_ARG(workspace_bytes >= weight_fit_workspace_bytes(utt_offsets_host[n_utt], k, n_utt), -2,
                "weight_fit: workspace too small");
  return launch_weight_fit(idx, single_table(synth, n_pool), dim, utt_offsets_host, n_utt, k, loss_scale, max_iters,
                           amp_ratio, out_weights, info, workspace, (cudaStream_t)stream);
}

int knnsvc_weight_fit_sharded(const int64_t* idx, const void* const* shard_rows_host, const int64_t* shard_lo_host,
                              int n_shards, int dim, const int64_t* utt_offsets_host, int n_utt, int k,
                              double loss_scale, int max_iters, float* out_weights, double* info, void* workspace,
                              size_t workspace_bytes, void* stream) {
  KNN_CHECK_ARG(idx && out_weights && workspace && utt_offsets_host && n_utt >= 0, -1, "weight_fit: bad arguments");
  if (n_utt == 0) return 0;
  KNN_CHECK_ARG(workspace_bytes >= weight_fit_workspace_bytes(utt_offsets_host[n_utt], k, n_utt), -2,
                "weight_fit: workspace too small");
  RowTable tab;
  int rc = make_table(shard_rows_host, shard_lo_host, n_shards, &tab, "weight_fit_sharded");
  if (rc) return rc;
  return launch_weight_fit(idx, tab, dim, utt_offsets_host, n_utt, k, loss_scale, max_iters, nullptr, out_weights, info,
                           workspace, (cudaStream_t)stream);
}

int knnsvc_harmonic_bank(const float* f0, const float* amp, int batch, int64_t frames, int n_harm, int sample_rate,
                         int hop, float* out, double* phase_ws, void* stream) {
  KNN_CHECK_ARG(f0 && out && phase_ws && batch >= 0 && frames >= 0, -1, "harmonic_bank: bad arguments");
  return launch_harmonic_bank(f0, amp, batch, frames, n_harm, sample_rate, hop, out, phase_ws, (cudaStream_t)stream);
}

int knnsvc_layer_mix(const float* feats, int n_layers, int64_t frames, int dim, const double* weights_a_host,
                     const double* weights_b_host, float* out_a, float* out_b, void* stream) {
  KNN_CHECK_ARG(feats && weights_a_host && out_a && frames >= 0 && dim >= 1, -1, "layer_mix: bad arguments");
  KNN_CHECK_ARG((weights_b_host == nullptr) == (out_b == nullptr), -1,
                "layer_mix: weights_b_host and out_b must be given together");
  return launch_layer_mix(feats, n_layers, frames, dim, weights_a_host, weights_b_host, out_a, out_b,
                          (cudaStream_t)stream);
}

int knnsvc_stft_magnitude(const float* audio, int64_t n_samples, int64_t frames, int n_fft, int hop, float* out,
                          void* stream) {
  KNN_CHECK_ARG(audio && out && n_samples >= 1 && frames >= 0, -1, "stft_magnitude: bad arguments");
  return launch_stft_magnitude(audio, n_samples, frames, n_fft, hop, out, (cudaStream_t)stream);
}

int knnsvc_harmonic_amplitudes(const float* spec, const float* f0, int64_t frames, int n_bins, int n_harm,
                               int sample_rate, float* out, void* stream) {
  KNN_CHECK_ARG(spec && f0 && out && frames >= 0 && sample_rate > 0, -1, "harmonic_amplitudes: bad arguments");
  return launch_harmonic_amplitudes(spec, f0, frames, n_bins, n_harm, sample_rate, out, (cudaStream_t)stream);
}

int knnsvc_row_l1(const float* x, int64_t rows, int dim, float* out, void* stream) {
  KNN_CHECK_ARG(x && out && rows >= 0 && dim >= 1, -1, "row_l1: bad arguments");
  return launch_row_l1(x, rows, dim, out, (cudaStream_t)stream);
}

int knnsvc_amp_ratio(const float* l1_query, const float* l1_pool, const int64_t* idx, int64_t n_query, int k,
                     int64_t n_pool, float* out, void* stream) {
  KNN_CHECK_ARG(l1_query && l1_pool && idx && out && n_query >= 0 && k >= 1 && n_pool >= 1, -1,
                "amp_ratio: bad arguments");
  return launch_amp_ratio(l1_query, l1_pool, idx, n_query, k, n_pool, out, (cudaStream_t)stream);
}

int knnsvc_store_to_host(const void* src_device, void* dst_pinned_host, size_t nbytes, void* stream) {
  KNN_CHECK_ARG(nbytes % 4 == 0, -1, "store_to_host: size must be a multiple of 4 bytes");
  if (nbytes == 0) return 0;
  KNN_CHECK_ARG(src_device && dst_pinned_host, -1, "store_to_host: null pointer");
  return launch_store_to_host(src_device, dst_pinned_host, nbytes, (cudaStream_t)stream);
}

}  // extern "C"
