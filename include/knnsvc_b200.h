/*
 * knnsvc_b200.h — C ABI of the B200-native kNN-SVC matcher hot path.
 *
 * The reference (SmoothKen/knn-svc) has no FFI layer: its seam is a set of
 * Python functions.  Each entry point below is the device-side
 * replacement for one of those functions; the Python mirror in
 * knn_svc_b200/ (same names and argument meaning as the reference) is a thin
 * ctypes caller of this library.  INTEGRATION.md shows the binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless it says "host";
 *   - the caller owns all memory, including workspaces (sizes are queried); the one exception
 *     is the small stream-ordered scratch of knnsvc_concat_cost_reselect (see there);
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*), on the
 *     current device; nothing synchronises the host;
 *   - return value 0 = ok, >0 = cudaError_t, <0 = argument error;
 *     knnsvc_last_error() returns a message for the calling thread;
 *   - there is NO CPU fallback anywhere behind this ABI.
 */
#ifndef KNNSVC_B200_H
#define KNNSVC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KNNSVC_MAX_K 32          /* ddsp_prematch_dataset.py:1203 takes k=32 */
#define KNNSVC_GREEDY_K 4        /* ddsp_prematch_dataset.py:1246 keeps 4    */

const char* knnsvc_last_error(void);
int knnsvc_version(void);

/* ---- K1 pre-pass: row norms + unit-normalised fp16 copy ------------------
 * Replaces torch.norm(x, p=2, dim=-1) at lib_ongaku_test.py:150-151 /
 * ddsp_matcher.py:215-216, and prepares the tensor-core operand.
 *   x        [rows, ld] fp32, first `dim` columns used
 *   half_out [rows, dim_pad] fp16 = fp16(x * 1024/|x|), zero in the pad columns
 *   norms    [rows] fp64 = |x| (fp64-accumulated; the re-score divides by these, so that its
 *            distances carry no fp32 rounding of the norms: ~1e-16, not 6e-8)
 *   bad_rows int counter, incremented per zero-norm / non-finite row (the
 *            reference NaNs and exits there, lib_ongaku_test.py:166-169)
 *   max_err  optional float, ZEROED BY THE CALLER: receives max over rows of
 *            |operand/2^10 - x/|x||_2, the measured rounding error the search's
 *            rigorous error window is built from (NULL: not measured; the search
 *            then assumes the fp16 worst case)
 */
int knnsvc_prepare_rows(const float* x, int64_t rows, int dim, int64_t ld,
                        void* half_out, int dim_pad, double* norms,
                        int* bad_rows, float* max_err, void* stream);

/* ---- K1 full matrix (API parity only; never on the fused path) -----------
 * fast_cosine_dist(source_feats, matching_pool) — lib_ongaku_test.py:148-175,
 * ddsp_matcher.py:213-221.  out [n_query, n_pool] fp32 = 1 - q.p/(|q||p|).
 */
int knnsvc_cosine_dist(const float* q, int64_t n_query, const float* p, int64_t n_pool,
                       int dim, float* out, void* stream);

/* ---- K1+K2 fused: cosine-distance kNN, k smallest per query row ----------
 * Replaces the chunk-20 loop fast_cosine_dist + .topk(k, largest=False) at
 * ddsp_prematch_dataset.py:1196-1206 and ddsp_matcher.py:550-554.
 * tcgen05 fp16 GEMM (fp32 accumulate in TMEM) filters candidates with a
 * rigorous error window, survivors are re-scored exactly from the fp32 rows;
 * rows the window cannot decide go through an exact brute-force kernel.
 * The [n_query, n_pool] matrix is never written.
 *   q/p       fp32 rows [n, dim] (ld = dim); qh/ph/qn/pn from knnsvc_prepare_rows
 *   index_offset  added to every returned index (pool shard offset, C1)
 *   q_err/p_err   the max_err scalars knnsvc_prepare_rows produced for the two row
 *             sets (device pointers; NULL = fp16 worst case, a wider window, same results)
 *   out_dist  [n_query, k] fp32 ascending;  out_idx [n_query, k] int64
 *   stats     optional int[8]: {rows through the exact fallback, logged candidates, candidates above
 *             the first (fp16-window) threshold = refined in fp32, segments, units, grid, log cap,
 *             candidates inside the refined window = scored in fp64}
 */
size_t knnsvc_knn_workspace_bytes(int64_t n_query, int64_t n_pool, int dim_pad, int k);
/* The traversal the search would use for this shape (host-only, no GPU work): plan_host int[8] <-
 * {CTAs per MMA, query tiles, pool tiles, pool segments, blocks per chain, work units, grid, log cap}. */
int knnsvc_knn_plan(int64_t n_query, int64_t n_pool, int k, int* plan_host);
int knnsvc_knn_search(const float* q, const void* qh, const double* qn, int64_t n_query,
                      const float* p, const void* ph, const double* pn, int64_t n_pool,
                      int dim, int dim_pad, int k, int64_t index_offset,
                      const float* q_err, const float* p_err,
                      float* out_dist, int64_t* out_idx,
                      void* workspace, size_t workspace_bytes, int* stats, void* stream);

/* Same search with a per-row masked column range: for query row t the pool columns
 * [mask_lo[t], mask_hi[t]) (shard-local, before index_offset) have their cosine distance
 * DEFINED as 1, exactly what the offline prematch does to an utterance's own frames —
 * `dists[:, start_index:end_index] = 1` before `.topk(k=32)`, ddsp_prematch_dataset.py:1608-1632
 * (per_spk_extract).  mask_lo/mask_hi: device int64 [n_query]; both NULL = knnsvc_knn_search. */
int knnsvc_knn_search_masked(const float* q, const void* qh, const double* qn, int64_t n_query,
                             const float* p, const void* ph, const double* pn, int64_t n_pool,
                             int dim, int dim_pad, int k, int64_t index_offset,
                             const float* q_err, const float* p_err,
                             const int64_t* mask_lo, const int64_t* mask_hi,
                             float* out_dist, int64_t* out_idx,
                             void* workspace, size_t workspace_bytes, int* stats, void* stream);

/* The general form of the search: the masked search plus `out_dist64` (optional, [n_query, k]):
 * the fp64 cosine distances the re-score ranked by (out_dist is their fp32 rounding).  The
 * sharded path (C1) exchanges and merges THESE, so a pool searched in N shards returns bit for bit
 * what one search of the whole pool returns. */
int knnsvc_knn_search_full(const float* q, const void* qh, const double* qn, int64_t n_query,
                           const float* p, const void* ph, const double* pn, int64_t n_pool,
                           int dim, int dim_pad, int k, int64_t index_offset,
                           const float* q_err, const float* p_err,
                           const int64_t* mask_lo, const int64_t* mask_hi,
                           float* out_dist, double* out_dist64, int64_t* out_idx,
                           void* workspace, size_t workspace_bytes, int* stats, void* stream);

/* Where the search keeps its candidate log inside the caller's workspace (host-only, no GPU work;
 * test / diagnostic aid — the parity tests read the tensor-core similarities s~ back from it):
 * layout_host int64[8] <- {byte offset of log_val (float [n_query*n_seg][cap]), of log_idx
 * (int32, same shape), of log_cnt (int32 [n_query*n_seg], cap+1 = overflowed), of seg_top
 * (float [n_query*n_seg][k]), n_seg, cap, total bytes, byte offset of ref_val (float, shape of
 * log_val: the fp32-refined similarity of each entry, -inf where s~ fell below the row's first
 * threshold)}. */
int knnsvc_knn_workspace_layout(int64_t n_query, int64_t n_pool, int dim_pad, int k, int64_t* layout_host);

/* Measurement hooks used by bench.py (no effect on results).
 *   knnsvc_launch_count            kernels this library has launched in this process
 *   knnsvc_filter_timing(1/0)      bracket the tcgen05 filter launch of every later
 *                                  knnsvc_knn_search with CUDA events on its stream
 *   knnsvc_filter_timing_collect   host float[max_n] <- per-call filter durations (ms)
 *                                  since the last collect; returns the count */
long long knnsvc_launch_count(void);
/* tuning / experiment switches: "cta_group" = 1 (the cta_group::2 filter variant of round 1 was removed),
 * "bf16_operands" = 0|1 (bf16 instead of fp16 tensor-core operands; measurement only,
 * the error window is sized for fp16), "concat_staged" = 1|0 (shared-memory staged K5
 * kernel where the row shape allows it, or the general kernel only), "concat_cluster" = 1|0 (K5: launches of
 * few utterances — as many as the device hosts clusters of 8 CTAs at once — run one utterance per cluster,
 * the feature dimension split over its CTAs; same results as the one-CTA staged kernel, bit for bit),
 * "concat_f0_table" = 1|0 (the cluster kernel reads log2 f0 of the pool rows from a table built once per
 * call, for pools up to 4M rows; same results),
 * "spin_sleep_ns" (barrier poll back-off of the filter's producer / MMA lanes),
 * "block_tiles" (pool tiles of 256 rows per L2 block of the filter traversal, 0 = default 96),
 * "query_group" (chains — query tile x segment — per group of the filter's two-level unit order,
 * 0 = default 2 x SM count; a value >= the number of chains gives the flat block-major order),
 * "refine_min_candidates" (candidates per row above the first threshold from which the decision stage
 * re-scores in fp32 before the fp64 decision, 0 = default 400; the route is chosen per query row; same results
 * either way),
 * "log_cap" (candidate-log slots per row and pool segment, 0 = default 2048; the tests shrink it to
 * drive rows into the overflow -> exact-kernel path with small fixtures),
 * "filter_flags" (bit0: L2 prefetch of the next unit's query tile [default on], bit2: static
 * instead of dynamic unit scheduling), "weight_fit_cluster" = 1|0 (K6: a cluster of 8 CTAs per
 * utterance for launches of few long utterances, or always one CTA per utterance; same results). */
int knnsvc_set_option(const char* name, int value);
int knnsvc_filter_timing(int enable);
int knnsvc_filter_timing_collect(float* ms_host, int max_n);

/* Exact brute-force kNN on CUDA cores (same outputs as knnsvc_knn_search).
 * Used for rows the filter flags, and by tests as an independent GPU check. */
size_t knnsvc_knn_exact_workspace_bytes(int64_t n_query, int64_t n_pool, int k);
int knnsvc_knn_exact(const float* q, const double* qn, int64_t n_query,
                     const float* p, const double* pn, int64_t n_pool, int dim, int k,
                     int64_t index_offset, float* out_dist, int64_t* out_idx,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ---- C1: merge per-shard top-k lists (after the NCCL all-gather) ---------
 * gathered_dist/idx [n_shards, n_query, k]; out [n_query, k]; ties -> lower index,
 * so the result does not depend on the shard count. */
int knnsvc_merge_topk(const float* gathered_dist, const int64_t* gathered_idx, int n_shards,
                      int64_t n_query, int k, float* out_dist, int64_t* out_idx, void* stream);

/* fp64 variant: gathered_dist is double [n_shards, n_query, k] (knnsvc_knn_search_full's
 * out_dist64); out_dist64 optional. */
int knnsvc_merge_topk64(const double* gathered_dist, const int64_t* gathered_idx, int n_shards,
                        int64_t n_query, int k, float* out_dist, double* out_dist64, int64_t* out_idx,
                        void* stream);

/* ---- C1: peer memory.  One process per GPU; a rank exports the allocation that holds its pool
 * shard's fp32 rows and opens its peers' (CUDA IPC, peer access enabled lazily), so that the
 * gather below can read matched rows straight from the GPU that owns them over NVLink.
 *   knnsvc_ipc_export  handle_host: 64 bytes <- cudaIpcMemHandle_t of the allocation `ptr` lies in;
 *                      offset_host <- byte offset of `ptr` inside that allocation
 *   knnsvc_ipc_open    base_out <- device pointer of the peer's allocation in THIS process
 *   knnsvc_ipc_close   unmaps it */
int knnsvc_ipc_export(const void* ptr, void* handle_host, int64_t* offset_host);
int knnsvc_ipc_open(const void* handle_host, void** base_out);
int knnsvc_ipc_close(void* base);

/* ---- K3 over a sharded pool: the gather + mix of knnsvc_gather_mix where pool rows
 * [shard_lo_host[s], shard_lo_host[s+1]) live at shard_rows_host[s] (this GPU's memory or an
 * IPC-mapped peer pointer).  idx holds GLOBAL row indices.  Same arithmetic, same order: the
 * result is bit-identical to mixing from one contiguous pool.  shard_* are HOST arrays of
 * n_shards (<= 16) pointers / n_shards + 1 bounds. */
int knnsvc_gather_mix_sharded(const void* const* shard_rows_host, const int64_t* shard_lo_host, int n_shards,
                              int dim, const int64_t* idx, const float* weights, int64_t n_query, int k,
                              float* out, void* stream);

/* ---- K5 / K6 over a sharded pool: the post-opt stage on a pool that lives in several row blocks
 * (same table as knnsvc_gather_mix_sharded; idx holds GLOBAL row indices).  The greedy re-selection
 * follows `previous selection + 1` (lib_ongaku_test.py:294-295) and the weight fit gathers rows
 * idx-1, idx, idx+1 (ddsp_prematch_dataset.py:585-590) ACROSS shard boundaries; rows of other GPUs are
 * fetched through their mapped pointers (TMA bulk copies / loads over NVLink).  pool_f0 is the
 * f0 of the WHOLE pool (replicated, n floats).  Results are bit-identical to the single-pool calls. */
int knnsvc_concat_cost_reselect_sharded(const int64_t* idx, const float* src,
                                        const void* const* shard_rows_host, const int64_t* shard_lo_host,
                                        int n_shards, int dim, const float* shifted_src_f0, const float* pool_f0,
                                        float concat_weight, const int64_t* utt_offsets_host, int n_utt,
                                        int64_t* out_idx, void* stream);
int knnsvc_weight_fit_sharded(const int64_t* idx, const void* const* shard_rows_host,
                              const int64_t* shard_lo_host, int n_shards, int dim,
                              const int64_t* utt_offsets_host, int n_utt, int k, double loss_scale, int max_iters,
                              float* out_weights, double* info, void* workspace, size_t workspace_bytes,
                              void* stream);

/* ---- K3: gather + weighted mix -------------------------------------------
 * out[t,:] = sum_k w[t,k] * pool[idx[t,k],:]  (w == NULL -> mean) —
 * ddsp_prematch_dataset.py:1348,1358,1364,1435,1444,1446; ddsp_matcher.py:578. */
int knnsvc_gather_mix(const float* pool, int64_t n_pool, int dim, const int64_t* idx,
                      const float* weights, int64_t n_query, int k, float* out, void* stream);

/* ---- K4: f0-compatibility re-rank ----------------------------------------
 * sort_by_f0_compatibility — ddsp_prematch_dataset.py:954-997: stable ascending
 * sort of each row's k candidates by |log2(f0[idx]+1e-5) - log2(expected+1e-5)|. */
int knnsvc_f0_rerank(const float* expected_f0, const float* pool_f0, const int64_t* idx,
                     int64_t n_query, int k, int64_t* out_idx, void* stream);

/* ---- K5: greedy concatenation-cost re-selection ---------------------------
 * knn_with_concat_cost — lib_ongaku_test.py:270-369, K = 4 candidates per frame.
 * Utterance u owns query rows [utt_offsets[u], utt_offsets[u+1]) (host array);
 * one CTA walks one utterance.  src_f0/pool_f0 NULL selects the no-f0 branch.
 * The only call that allocates: (n_utt + 1) offsets + 2 doubles per frame of stream-ordered
 * scratch from a private cudaMemPool (kept across synchronisations, freed on the stream). */
int knnsvc_concat_cost_reselect(const int64_t* idx, const float* src, const float* pool,
                                int64_t n_pool, int dim, const float* shifted_src_f0,
                                const float* pool_f0, float concat_weight,
                                const int64_t* utt_offsets_host, int n_utt,
                                int64_t* out_idx, void* stream);

/* ---- K6: Adam(amsgrad) fit of per-frame softmax mixing weights ------------
 * compute_wavlm_weight (loss_scale 0.1) — ddsp_prematch_dataset.py:574-680;
 * compute_extended_weight (loss_scale 1000) — :807-924.
 *   info (device, optional) double[4]: {stop iteration, best loss, first loss, 0} */
size_t knnsvc_weight_fit_workspace_bytes(int64_t n_query, int k);
int knnsvc_weight_fit(const int64_t* idx, const float* synth, int64_t n_pool, int dim,
                      int64_t n_query, int k, double loss_scale, int max_iters,
                      float* out_weights, double* info,
                      void* workspace, size_t workspace_bytes, void* stream);

/* Batched form (BASELINE cfg 5: many utterances against one pool): utterance u owns frames
 * [utt_offsets_host[u], utt_offsets_host[u+1]) of idx / out_weights; one CTA per utterance,
 * every utterance with its own stop rule.  info (optional): double[n_utt][4]. */
size_t knnsvc_weight_fit_batched_workspace_bytes(int64_t n_frames, int k, int n_utt);
int knnsvc_weight_fit_batched(const int64_t* idx, const float* synth, int64_t n_pool, int dim,
                              const int64_t* utt_offsets_host, int n_utt, int k, double loss_scale,
                              int max_iters, float* out_weights, double* info,
                              void* workspace, size_t workspace_bytes, void* stream);

/* Training-time variant (offline prematch): compute_weight_with_amp —
 * ddsp_prematch_dataset.py:684-803, called at :1681 with loss 1000*MSE (phase_mae :449-457).
 * Every candidate row is scaled by amp_ratio[t,k] (device fp32 [n_frames, k]) before mixing;
 * amp_ratio == NULL is knnsvc_weight_fit_batched. */
int knnsvc_weight_fit_amp(const int64_t* idx, const float* synth, int64_t n_pool, int dim,
                          const int64_t* utt_offsets_host, int n_utt, int k, double loss_scale,
                          int max_iters, const float* amp_ratio, float* out_weights, double* info,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ---- K7 / K7': additive harmonic bank -------------------------------------
 * get_bulk_dsp_choral — ddsp_prematch_dataset.py:165-208 (amp != NULL, H harmonics)
 * and the single sinusoid of hifigan/ddsp_models_f0.py:344-352 (amp == NULL).
 *   f0 [batch, frames]; amp [batch, frames, n_harm]; out [batch, frames*hop] fp32
 *   phase_ws: double[batch*frames] scratch */
int knnsvc_harmonic_bank(const float* f0, const float* amp, int batch, int64_t frames,
                         int n_harm, int sample_rate, int hop, float* out,
                         double* phase_ws, void* stream);

/* ==== SURVEY §8(f): the tensor ops either side of the matcher ==============
 * Pool builder (get_complete_spk_pool, ddsp_prematch_dataset.py:301-423) after WavLM: */

/* (feats*weights[:, None]).sum(dim=0) — :349-350.  feats [n_layers, frames, dim] fp32;
 * weights_*_host: HOST double[n_layers] (the reference's weighting is float64);
 * both mixes (matching + synthesis weights) in one pass; weights_b_host/out_b may be NULL. */
int knnsvc_layer_mix(const float* feats, int n_layers, int64_t frames, int dim,
                     const double* weights_a_host, const double* weights_b_host,
                     float* out_a, float* out_b, void* stream);

/* torchaudio Spectrogram(n_fft=400, hop_length=320, center=True, power=1)(x).T[:, :-1][:frames]
 * — :326, :361-363.  audio [n_samples] fp32 -> out [frames, n_fft/2] fp32 magnitudes. */
int knnsvc_stft_magnitude(const float* audio, int64_t n_samples, int64_t frames, int n_fft, int hop,
                          float* out, void* stream);

/* Harmonic amplitudes read off the x8-interpolated spectrum — :391-404.
 * spec [frames, n_bins] (n_bins = 200), f0 [frames] -> out [frames, n_harm] (n_harm = 49). */
int knnsvc_harmonic_amplitudes(const float* spec, const float* f0, int64_t frames, int n_bins,
                               int n_harm, int sample_rate, float* out, void* stream);

/* Offline prematch (per_spk_extract :1672-1675): amp_ratio[t,k] =
 * |spec_utt[t]|_1 / (|spec_pool[idx[t,k]]|_1 + 1e-5).  knnsvc_row_l1: out[r] = sum |x[r,:]|. */
int knnsvc_row_l1(const float* x, int64_t rows, int dim, float* out, void* stream);
/* Small result -> PINNED host memory by a kernel (stores over PCIe into the mapped allocation; with unified
 * addressing every cudaHostAlloc / torch pinned allocation is device-accessible at its host address) instead of
 * a cudaMemcpy: a blocking device-to-host read waits on the copy engine behind whatever large download another
 * stream has queued there, this does not.  Extension (no reference counterpart): it is how the matcher hands
 * the shifted f0 back on the host (ddsp_prematch_dataset.py:1224-1233 computes it there) while a previous
 * batch's features are still being downloaded.  nbytes a multiple of 4; the caller synchronises the stream. */
int knnsvc_store_to_host(const void* src_device, void* dst_pinned_host, size_t nbytes, void* stream);

int knnsvc_amp_ratio(const float* l1_query, const float* l1_pool, const int64_t* idx,
                     int64_t n_query, int k, int64_t n_pool, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KNNSVC_B200_H */
