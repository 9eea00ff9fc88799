// Internal launcher declarations shared by the .cu files and the C ABI (capi.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace knnsvc {

// ---- tuning options (capi.cu): diagnostic switches, see knnsvc_set_option
int opt_bf16();
int opt_filter_flags();   // bit0: L2 prefetch of the next unit's query tile, bit2: static unit schedule (bit1 unused)
int opt_weight_fit_cluster();   // 1 (default): few long utterances are fitted by a cluster of 8 CTAs each
int opt_refine_min();      // candidates per row above which the decision stage refines in fp32 first (0 = default)
int opt_log_cap();       // candidate-log slots per (row, segment); 0 = default
int opt_query_group();   // chains per group of the filter's two-level unit order (0 = default)
int opt_block_tiles();   // pool tiles per L2 block of the filter traversal (0 = default)
int opt_concat_staged();   // 1 (default): shared-memory staged K5 where eligible; 0: general kernel only
int opt_concat_cluster();  // 1 (default): a few long utterances get a cluster of 8 CTAs each (dimension split)
int opt_concat_f0_table(); // 1 (default): the cluster kernel reads log2 f0 of pool rows from a per-call table (pools <= 4M rows)
int opt_epi_sleep_ns();  // nanosleep between the epilogue warps' polls of the accumulator-ready barrier
int opt_spin_ns();     // nanosleep between barrier polls of the producer / MMA lanes (0 = pure spin)        // 1: bf16 tensor-core operands (experiment only: 8x wider rounding error than fp16)

// ---- rows.cu
int launch_prepare_rows(const float* x, int64_t rows, int dim, int64_t ld, void* half_out, int dim_pad,
                        double* norms, int* bad_rows, float* max_err, cudaStream_t stream);
int launch_cosine_dist(const float* q, int64_t nq, const float* p, int64_t np_, int dim, float* out,
                       cudaStream_t stream);
int exact_chunks(int64_t n_pool);
size_t exact_partial_bytes(int64_t slots, int64_t n_pool, int k);
int launch_knn_exact_rows(const float* q, const double* qn, int64_t n_query, const float* p, const double* pn,
                          int64_t n_pool, int dim, int k, const int64_t* row_list, const int* row_count_dev,
                          int64_t row_count_host, int64_t slot_base, int64_t slot_cap, int64_t index_offset,
                          float* out_dist, double* out_dist64, int64_t* out_idx, void* partial,
                          const int64_t* mask_lo, const int64_t* mask_hi, cudaStream_t stream);

// ---- knn_filter_sm100.cu
struct FilterPlan {
  int ctas, n_qtiles, n_ptiles, n_seg, n_blk, n_units, grid, cap;
  int rf_rows, rf_nblk;   // refine blocks (knn_select.cu): pool rows per block, number of blocks
  int grp;                // chains per group of the filter's two-level unit order
};
FilterPlan plan_filter(int64_t n_query, int64_t n_pool, int k);
void plan_refine(int64_t n_query, int64_t n_pool, int dim, int* rf_rows, int* rf_nblk);
int launch_knn_filter(const void* qh, int64_t n_query, const void* ph, int64_t n_pool, int dim_pad, int k,
                      const FilterPlan& pl, float* log_val, int* log_idx, int* log_cnt, float* seg_top,
                      float* seg_kth, int* seg_flag, const int64_t* mask_lo, const int64_t* mask_hi,
                      const float* q_err, const float* p_err, int* unit_counter, cudaStream_t stream);
size_t filter_flag_count(const FilterPlan& pl);

// ---- knn_select.cu
constexpr int kFlagCap = 1024;  // rows the in-call exact fallback can absorb
int launch_knn_rescore(const float* q, const double* qn, int64_t n_query, const float* p, const double* pn,
                       int64_t n_pool, int dim, int k, const FilterPlan& pl, const float* log_val,
                       const int* log_idx, const int* log_cnt, const float* seg_top, float* ref_val, float* row_thr,
                       int* row_mode, int* blk_off, int64_t index_offset, float* out_dist, double* out_dist64, int64_t* out_idx,
                       int64_t* flag_list, int* flag_count, int* stats, const int64_t* mask_lo, const int64_t* mask_hi,
                       const float* q_err, const float* p_err, cudaStream_t stream);
int launch_merge_topk(const float* gd, const int64_t* gi, int n_shards, int64_t n_query, int k, float* out_dist,
                      int64_t* out_idx, cudaStream_t stream);
int launch_merge_topk64(const double* gd, const int64_t* gi, int n_shards, int64_t n_query, int k, float* out_dist,
                        double* out_dist64, int64_t* out_idx, cudaStream_t stream);

// ---- post.cu
// A pool stored as up to kMaxShards row blocks: rows [lo[s], lo[s+1]) live at base[s] (this GPU's
// memory or a peer's, mapped through CUDA IPC — loads then travel over NVLink).
constexpr int kMaxShards = 16;
struct RowTable {
  const float* base[kMaxShards];
  int64_t lo[kMaxShards + 1];
  int n;
};
inline RowTable single_table(const float* rows, int64_t n_rows) {
  RowTable t;
  t.n = 1;
  t.base[0] = rows;
  t.lo[0] = 0;
  t.lo[1] = n_rows;
  return t;
}
#ifdef __CUDACC__
// pointer to global row r (clamped into the table's range, as the reference clamps `prev + 1` and
// `idx +- 1` at the ends of the pool: lib_ongaku_test.py:294-295, ddsp_prematch_dataset.py:585-590)
__device__ __forceinline__ const float* table_row(const RowTable& tab, int64_t r, int dim) {
  r = r < 0 ? 0 : (r >= tab.lo[tab.n] ? tab.lo[tab.n] - 1 : r);
  int s = 0;
#pragma unroll 1
  while (s + 1 < tab.n && r >= tab.lo[s + 1]) ++s;
  return tab.base[s] + (r - tab.lo[s]) * dim;
}
__device__ __forceinline__ bool table_aligned16(const RowTable& tab) {
  bool ok = true;
  for (int s = 0; s < tab.n; ++s) ok = ok && ((reinterpret_cast<uintptr_t>(tab.base[s]) & 15) == 0);
  return ok;
}
#endif
int launch_gather_mix_sharded(const RowTable& tab, int dim, const int64_t* idx, const float* weights, int64_t n_query,
                              int k, float* out, cudaStream_t stream);
int launch_gather_mix(const float* pool, int64_t n_pool, int dim, const int64_t* idx, const float* weights,
                      int64_t n_query, int k, float* out, cudaStream_t stream);
int launch_f0_rerank(const float* expected_f0, const float* pool_f0, const int64_t* idx, int64_t n_query, int k,
                     int64_t* out_idx, cudaStream_t stream);
int launch_concat_cost(const int64_t* idx, const float* src, const RowTable& pool, int dim,
                       const float* src_f0, const float* pool_f0, float concat_weight, const int64_t* utt_offsets_dev,
                       int n_utt, int64_t n_frames, double* frame_ws, double* lf0_ws, int64_t* out_idx,
                       cudaStream_t stream);   // lf0_ws: optional scratch of one double per pool frame (f0 runs of few utterances)

// ---- concat_cost_sm100.cu
bool concat_staged_eligible(const float* src, const RowTable& pool, int dim);
int launch_concat_cost_staged(const int64_t* idx, const float* src, const RowTable& pool, int dim,
                              const float* src_f0, const float* pool_f0, float concat_weight,
                              const int64_t* utt_offsets_dev, int n_utt, const double* base, const double* n2,
                              int64_t* out_idx, cudaStream_t stream);

bool concat_cluster_fits(int n_utt);   // true: every utterance gets a resident cluster of 8 CTAs on this device
int launch_concat_cost_cluster(const int64_t* idx, const float* src, const RowTable& pool, int dim,
                               const float* src_f0, const float* pool_f0, const double* lf0_tab, float concat_weight,
                               const int64_t* utt_offsets_dev, int n_utt, const double* base, const double* n2,
                               int64_t* out_idx, cudaStream_t stream);

int launch_store_to_host(const void* src, void* dst_mapped, size_t nbytes, cudaStream_t stream);   // post.cu

// ---- weight_fit.cu
size_t weight_fit_workspace_bytes(int64_t n_query, int k, int n_utt);
int launch_weight_fit(const int64_t* idx, const RowTable& synth, int dim,
                      const int64_t* utt_offsets_host, int n_utt, int k, double loss_scale, int max_iters,
                      const float* amp, float* out_weights, double* info, void* workspace, cudaStream_t stream);

// ---- harmonic.cu
int launch_harmonic_bank(const float* f0, const float* amp, int batch, int64_t frames, int n_harm, int sample_rate,
                         int hop, float* out, double* phase_ws, cudaStream_t stream);

// ---- pool_ops.cu
constexpr int kMaxLayers = 32;   // WavLM-Large exposes 25 layer outputs
int launch_layer_mix(const float* feats, int n_layers, int64_t frames, int dim, const double* w_a_host,
                     const double* w_b_host, float* out_a, float* out_b, cudaStream_t stream);
int launch_stft_magnitude(const float* audio, int64_t n_samples, int64_t frames, int n_fft, int hop, float* out,
                          cudaStream_t stream);
int launch_harmonic_amplitudes(const float* spec, const float* f0, int64_t frames, int n_bins, int n_harm,
                               int sample_rate, float* out, cudaStream_t stream);
int launch_row_l1(const float* x, int64_t rows, int dim, float* out, cudaStream_t stream);
int launch_amp_ratio(const float* l1_query, const float* l1_pool, const int64_t* idx, int64_t n_query, int k,
                     int64_t n_pool, float* out, cudaStream_t stream);

}  // namespace knnsvc
