// K1+K2 fused: cosine-similarity GEMM on the 5th-gen tensor cores with a
// streaming candidate filter in the epilogue.  Replaces the reference's chunk-20
// loop  fast_cosine_dist(...) ; dists.topk(k=32, largest=False)
// (ddsp_prematch_dataset.py:1196-1206, lib_ongaku_test.py:148-175) without ever
// writing the [T, Np] distance matrix.
//
// Shape of the computation (persistent over work units):
//   one CTA per SM, UMMA M = 128, N = 256, A 16 KB + B 32 KB per stage, 4 stages.
//   (A cta_group::2 variant — CTA pair, UMMA M = 256, half of B per CTA — was built in round 1 and
//   measured again in round 2 at the 8-GPU shard size: 1.5-1.8% slower than this shape,
//   profiles/r2_cta_group_ab_1p25M.jsonl; a pair runs at the pace of its slower SM.  It was removed.)
//   D = 128 lanes x 256 fp32 columns per CTA in TMEM, double buffered (512 columns)
//   warp 0 : TMA producer      warp 1 : tcgen05.mma issuer + TMEM owner
//   warps 2-5 : epilogue, one thread per query row (TMEM lane), tcgen05.ld 32x32b
//   Operands: TMA, SWIZZLE_128B, K-major.
//
// A CHAIN is (query tile, pool segment): its 128 rows walk the segment's pool tiles with one
// running threshold per row.  A chain is cut into BLOCKS of ~96 pool tiles (an L2-sized piece of
// the fp16 pool) and a work unit is (chain, block).  Units are ordered block-major and CLAIMED
// DYNAMICALLY from a global counter (the producer lane claims, a 4-deep shared-memory queue hands
// the unit to the MMA lane and the epilogue warps): SMs differ by up to ~14% in sustained tensor
// rate under the power cap, so with a static split the fast ones idled at the end and the CTAs
// drifted GBs apart in the pool (it was re-read ~50x per launch); now a faster SM simply takes
// more units and all CTAs stay on the same pool block.  Inside a unit each epilogue thread carries
// its row's state in registers; between the blocks of a chain the state (top-k list, log count)
// goes through global memory, ordered by a per-warp block counter (consecutive blocks of a chain
// may run on different SMs, never concurrently).
//
// Filter rule (rigorous, DESIGN.md "exact top-k from an fp16 GEMM"): operands are
// unit-normalised rows cast to fp16, so the accumulator is the cosine similarity
// s~ with |s~ - s| <= eps (eps from the rows' MEASURED rounding errors, common.cuh).  The thread tracks tau_k = k-th largest s~ seen so far
// and LOGS every column with s~ > tau_k - 2*eps.  Every true top-k member is in
// the log; knn_rescore re-scores the log exactly from the fp32 rows.
#include <cuda.h>
#include <limits.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.cuh"

namespace knnsvc {

namespace {

constexpr int BM = 128;   // query rows per CTA (TMEM lanes)
constexpr int BN = 256;   // pool rows per MMA tile (TMEM columns)
constexpr int BK = 64;    // fp16 elements per pipeline stage along the feature dimension
constexpr int UMMA_K = 16;
constexpr int A_BYTES = BM * BK * 2;  // 16 KB
constexpr int NUM_THREADS = 192;
constexpr int TMEM_COLS = 512;
constexpr int kSchedDepth = 4;   // units the producer may run ahead of the slowest consumer warp

struct Cfg {
  static constexpr int B_ROWS = BN;                    // pool rows staged per tile
  static constexpr int B_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = 4;
  // operand ring first: SWIZZLE_128B wants 1024-byte aligned stage bases
  static constexpr int ring = 0;
  static constexpr int topv = ring + STAGES * STAGE_BYTES;              // float [kMaxK][BM]
  static constexpr int full_bar = topv + kMaxK * BM * 4;                // [STAGES]
  static constexpr int empty_bar = full_bar + STAGES * 8;               // [STAGES]
  static constexpr int tmem_full_bar = empty_bar + STAGES * 8;          // [2]
  static constexpr int tmem_empty_bar = tmem_full_bar + 2 * 8;          // [2]
  static constexpr int tmem_ptr = tmem_empty_bar + 2 * 8;               // uint32
  static constexpr int sched_full_bar = tmem_ptr + 16;                  // [kSchedDepth]  unit queue (dynamic scheduling)
  static constexpr int sched_empty_bar = sched_full_bar + kSchedDepth * 8;
  static constexpr int sched_unit = sched_empty_bar + kSchedDepth * 8;  // int [kSchedDepth]
  static constexpr int total = sched_unit + kSchedDepth * 4;
  static constexpr int SMEM_BYTES = total + 1024;  // slack for manual 1024 B alignment
};

// ----------------------------------------------------------------------------- PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
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
// wait used by the single-lane producer / MMA loops: optional nanosleep between polls
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity, uint32_t sleep_ns) {
  if (sleep_ns == 0) {
    mbar_wait(bar, parity);
    return;
  }
  uint32_t done;
  while (true) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(sleep_ns);
  }
}
// TMA tile load; the completion bytes are credited to `bar`
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// L2 prefetch of one TMA box (no shared-memory destination, no completion to wait for)
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// completion of all prior MMAs of this thread -> one arrival on `bar`
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane (base+i).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor: 8-row core groups of
// 1024 B (SBO), version 1 (Blackwell), layout type 2 (128-byte swizzle).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);        // start address, 16-byte units
  d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                          // descriptor version
  d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=f16, both K-major, N=256, M=128.
struct InstrDesc {
  static constexpr uint32_t value = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
};

struct RowState {
  float tau_lo;  // scaled: log everything strictly above this
  float kth;     // scaled lower bound of the k-th largest value seen (kEmpty until k values seen)
  int kpos;      // slot holding `kth`
  int cnt;       // entries in this row's log (cap+1 = overflowed)
};

constexpr float kEmpty = -3.0e38f;   // "no value yet" (finite, so the packed key decodes cleanly)
constexpr int kSlots = kMaxK;        // list slots scanned per update (slots >= k hold INT_MAX)

// Order-preserving map between fp32 bit patterns and signed ints (an involution).
__device__ __forceinline__ int mono(int b) { return b ^ ((b >> 31) & 0x7fffffff); }
// Key of a list entry: the value rounded DOWN to 27 significant bits with the slot id in the low
// 5 bits, so one min-reduction over the keys yields both the k-th value and where it lives.
// Rounding down keeps the tracked k-th value a lower bound of the true one (the filter stays a
// superset); the loss is < 4e-6 of the value, against an error window of 2.4e-3.
__device__ __forceinline__ int pack_key(float v, int slot) { return (mono(__float_as_int(v)) & ~31) | slot; }
__device__ __forceinline__ float key_value(int key) { return __int_as_float(mono(key & ~31)); }

// r[j] for a runtime j without local memory: a 5-level select tree (31 SELs).
__device__ __forceinline__ uint32_t select32(const uint32_t (&r)[32], int j) {
  uint32_t a[16], b[8], c[4], d[2];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (j & 1) ? r[2 * i + 1] : r[2 * i];
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = (j & 2) ? a[2 * i + 1] : a[2 * i];
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i] = (j & 4) ? b[2 * i + 1] : b[2 * i];
#pragma unroll
  for (int i = 0; i < 2; ++i) d[i] = (j & 8) ? c[2 * i + 1] : c[2 * i];
  return (j & 16) ? d[1] : d[0];
}

// Rare path, part 1 (slow): the row's log is full.  Compact it in place — entries below the
// current threshold can never be needed (tau only rises) — then append, or mark a genuine
// overflow (more than `cap` candidates inside the window: the exact kernel will decide the row).
__device__ __noinline__ RowState filter_log_full(RowState st, float v, int col, float* __restrict__ log_val,
                                                 int* __restrict__ log_idx, int cap) {
  if (st.cnt == cap) {
    int n = 0;
    for (int e0 = 0; e0 < cap; e0 += 8) {
      float lv[8];
      int li[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        lv[u] = __ldcg(log_val + e0 + u);   // L2: an earlier block of the chain may have run on another SM
        li[u] = __ldcg(log_idx + e0 + u);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (lv[u] * kDotScale > st.tau_lo) {
          log_val[n] = lv[u];
          log_idx[n] = li[u];
          ++n;
        }
      }
    }
    st.cnt = n;
  }
  if (st.cnt < cap) {
    log_val[st.cnt] = v * kDotUnscale;
    log_idx[st.cnt] = col;
    ++st.cnt;
  } else {
    st.cnt = cap + 1;  // genuine overflow
  }
  return st;
}

// Rare path, part 2: the value also beats the running k-th best.  It replaces that entry of the
// row's UNSORTED top-k list in shared memory ([slot][row]: the 32 rows of a warp hit 32 banks)
// and the minimum over the 32 slot keys is retaken — independent loads and a min tree instead of
// an insertion sort's dependent chain.
__device__ __noinline__ RowState filter_list_update(RowState st, float v, int* __restrict__ keys_row,
                                                    float window_scaled) {
  keys_row[st.kpos * BM] = pack_key(v, st.kpos);
  int key[kSlots];
#pragma unroll
  for (int j = 0; j < kSlots; ++j) key[j] = keys_row[j * BM];
#pragma unroll
  for (int w = kSlots / 2; w > 0; w >>= 1)
#pragma unroll
    for (int j = 0; j < w; ++j) key[j] = min(key[j], key[j + w]);
  st.kth = key_value(key[0]);
  st.kpos = key[0] & 31;
  st.tau_lo = st.kth - window_scaled;  // stays ~-3e38 until k values have been seen
  return st;
}

// One accumulator value passed the register threshold: log it (inline fast path — on rows with
// dense neighbourhoods most passing values sit inside the error window but below the k-th best and
// need nothing else) and update the top-k list only if it beats the k-th best.
__device__ __forceinline__ RowState filter_insert(RowState st, float v, int col, int* __restrict__ keys_row, int k,
                                                  float* __restrict__ log_val, int* __restrict__ log_idx, int cap,
                                                  float window_scaled) {
  (void)k;
  if (st.cnt < cap) {
    log_val[st.cnt] = v * kDotUnscale;
    log_idx[st.cnt] = col;
    ++st.cnt;
  } else {
    st = filter_log_full(st, v, col, log_val, log_idx, cap);
  }
  if (v > st.kth) st = filter_list_update(st, v, keys_row, window_scaled);
  return st;
}

// Work unit u -> (block, segment, query tile) and the unit's pool-tile range.  The chains (query tile x
// segment) are taken in GROUPS of `grp` chains; inside a group the order is block-major (all the group's
// chains' block 0, then block 1, ...), and the groups follow one another.  With ONE group (the default,
// see plan_filter) this is the flat block-major order; smaller groups were meant to keep a group's query
// tiles in L2 while the pool blocks stream past them — measured, they do not (plan_filter).  A unit only
// ever depends on the same chain's previous block, which has a lower unit index in every order.
struct Unit {
  int qt, seg, blk, t0, t1;
};
__device__ __forceinline__ Unit decode_unit(int u, int n_qtiles, int n_seg, int n_blk, int n_ptiles, int grp) {
  const int n_chains = n_qtiles * n_seg;
  const int per_group = grp * n_blk;
  const int g = u / per_group;
  const int r = u - g * per_group;
  const int c0 = g * grp;
  const int in_group = n_chains - c0 < grp ? n_chains - c0 : grp;
  Unit w;
  w.blk = r / in_group;
  const int c = c0 + (r - w.blk * in_group);
  w.seg = c / n_qtiles;
  w.qt = c - w.seg * n_qtiles;
  const int s0 = (int)((int64_t)w.seg * n_ptiles / n_seg), s1 = (int)((int64_t)(w.seg + 1) * n_ptiles / n_seg);
  w.t0 = s0 + (int)((int64_t)(s1 - s0) * w.blk / n_blk);
  w.t1 = s0 + (int)((int64_t)(s1 - s0) * (w.blk + 1) / n_blk);
  return w;
}

}  // namespace

// ----------------------------------------------------------------------------- kernel
// MASKED: every query row carries a column range [mask_lo[row], mask_hi[row]) whose cosine
// distance is DEFINED to be 1 (similarity 0) — the offline prematch's self-utterance rule
// `dists[:, start_index:end_index] = 1` (ddsp_prematch_dataset.py:1623-1624).  The accumulator
// values of those columns are replaced by 0 before the filter sees them.
template <bool MASKED>
__global__ void __launch_bounds__(NUM_THREADS, 1)
knn_filter_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_p,
                  int64_t n_query, int64_t n_pool, int k_blocks, int k, int n_qtiles, int n_ptiles, int n_seg,
                  int n_blk, int cap, float* __restrict__ log_val, int* __restrict__ log_idx, int* __restrict__ log_cnt,
                  float* __restrict__ seg_top, float* __restrict__ seg_kth, int* __restrict__ seg_flag,
                  uint32_t idesc, uint32_t spin_ns, const int64_t* __restrict__ mask_lo,
                  const int64_t* __restrict__ mask_hi, const float* __restrict__ q_err,
                  const float* __restrict__ p_err, int* __restrict__ unit_counter, int prefetch, int grp) {
  using L = Cfg;
  extern __shared__ unsigned char smem_raw_unaligned[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw_unaligned) + 1023) &
                                                         ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_units = n_qtiles * n_seg * n_blk;
  const int worker = blockIdx.x;               // index of this CTA among the persistent workers
  const int n_workers = gridDim.x;
  // dynamic unit scheduling: units are claimed from a global counter in index order (bit 2: static split)
  const bool dyn = !(prefetch & 4);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_p) : "memory");
    for (int s = 0; s < L::STAGES; ++s) {
      mbar_init(sbase + L::full_bar + s * 8, 1);
      mbar_init(sbase + L::empty_bar + s * 8, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(sbase + L::tmem_full_bar + b * 8, 1);
      mbar_init(sbase + L::tmem_empty_bar + b * 8, 4);  // one arrive per epilogue warp
    }
    for (int i = 0; i < kSchedDepth; ++i) {
      mbar_init(sbase + L::sched_full_bar + i * 8, 1);
      mbar_init(sbase + L::sched_empty_bar + i * 8, 5);   // the MMA lane + four epilogue warps
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + L::tmem_ptr),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + L::tmem_ptr);

  if (warp == 0) {
    // ===================== TMA producer (one per CTA; each stages its own A rows and its part of B) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0, sq = 0;
      uint32_t sq_phase = 0;
      // next unit of this worker: claimed from the global counter (dynamic: a faster SM simply takes
      // more units and all CTAs stay on the same pool block) or the static stride
      auto claim = [&]() -> int {
        int u;
        if (dyn) u = atomicAdd(unit_counter, 1);
        else u = worker + (it++) * n_workers;
        return u < n_units ? u : -1;
      };
      int u_next = claim();
      for (;;) {
        const int u = u_next;
        if (dyn) {   // hand the unit to the MMA lane and the epilogue warps (-1 = no more work)
          mbar_wait(sbase + L::sched_empty_bar + sq * 8, sq_phase ^ 1);
          *reinterpret_cast<volatile int*>(smem + L::sched_unit + sq * 4) = u;
          mbar_arrive(sbase + L::sched_full_bar + sq * 8);
          if (++sq == kSchedDepth) {
            sq = 0;
            sq_phase ^= 1;
          }
        }
        if (u < 0) break;
        const Unit un = decode_unit(u, n_qtiles, n_seg, n_blk, n_ptiles, grp);
        const int q_row = un.qt * BM;
        const int n_t = un.t1 - un.t0;
        // The next unit is claimed four tiles before this one ends (not at its start: CTAs that
        // launch first would grab second units before the last CTAs have taken their first), and
        // its query tile is prefetched into L2 over those tiles, a few k-slab boxes per tile —
        // otherwise every unit would start with 16 dependent DRAM round trips.
        const int a_span = n_t < 4 ? n_t : 4, a_first = n_t - a_span;
        int nq_row = -1;
        bool claimed = false;
        for (int pt = un.t0; pt < un.t1; ++pt) {
          const int ti = pt - un.t0;
          if (ti == a_first) {
            u_next = claim();
            claimed = true;
            if ((prefetch & 1) && u_next >= 0) {
              const Unit nx = decode_unit(u_next, n_qtiles, n_seg, n_blk, n_ptiles, grp);
              if (nx.qt != un.qt) nq_row = nx.qt * BM;
            }
          }
          if (nq_row >= 0) {
            const int j = ti - a_first;
            for (int kb = j * k_blocks / a_span; kb < (j + 1) * k_blocks / a_span; ++kb)
              tma_prefetch_2d(&map_q, kb * BK, nq_row);
          }
          const int p_row = pt * BN;
          for (int kb = 0; kb < k_blocks; ++kb) {
            mbar_wait_backoff(sbase + L::empty_bar + stage * 8, phase ^ 1, spin_ns & 0xffffu);
            const uint32_t full = sbase + L::full_bar + stage * 8;
            mbar_expect_tx(full, L::STAGE_BYTES);
            const uint32_t a_dst = sbase + L::ring + stage * L::STAGE_BYTES;
            tma_load_2d(a_dst, &map_q, full, kb * BK, q_row);
            tma_load_2d(a_dst + A_BYTES, &map_p, full, kb * BK, p_row);
            if (++stage == L::STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
        if (!claimed) u_next = claim();   // empty unit
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t tile_n = 0;
      int it = 0, sq = 0;
      uint32_t sq_phase = 0;
      for (;;) {
        int u;
        if (dyn) {
          mbar_wait(sbase + L::sched_full_bar + sq * 8, sq_phase);
          u = *reinterpret_cast<volatile int*>(smem + L::sched_unit + sq * 4);
          mbar_arrive(sbase + L::sched_empty_bar + sq * 8);
          if (++sq == kSchedDepth) {
            sq = 0;
            sq_phase ^= 1;
          }
        } else {
          u = worker + (it++) * n_workers;
          if (u >= n_units) u = -1;
        }
        if (u < 0) break;
        const Unit un = decode_unit(u, n_qtiles, n_seg, n_blk, n_ptiles, grp);
        for (int pt = un.t0; pt < un.t1; ++pt, ++tile_n) {
          const uint32_t buf = tile_n & 1;
          const uint32_t buf_phase = (tile_n >> 1) & 1;
          mbar_wait_backoff(sbase + L::tmem_empty_bar + buf * 8, buf_phase ^ 1, spin_ns & 0xffffu);
          tcgen05_fence_after();
          const uint32_t d_tmem = tmem_base + buf * BN;
          for (int kb = 0; kb < k_blocks; ++kb) {
            mbar_wait_backoff(sbase + L::full_bar + stage * 8, phase, spin_ns & 0xffffu);
            tcgen05_fence_after();
            const uint32_t a_addr = sbase + L::ring + stage * L::STAGE_BYTES;
            const uint64_t da = make_smem_desc(a_addr);
            const uint64_t db = make_smem_desc(a_addr + A_BYTES);
#pragma unroll
            for (int kk = 0; kk < BK / UMMA_K; ++kk) {
              // advancing 16 fp16 = 32 bytes inside the 128-byte swizzle row: +2 in 16-byte units
              umma_f16(d_tmem, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc, (kb | kk) != 0);
            }
            tcgen05_commit(sbase + L::empty_bar + stage * 8);  // smem slot free once these MMAs retire
            if (++stage == L::STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
          tcgen05_commit(sbase + L::tmem_full_bar + buf * 8);  // accumulator ready for the epilogue(s)
        }
      }
    }
  } else {
    // ===================== epilogue: streaming candidate filter =====================
    const int quad = warp & 3;                   // TMEM lane quadrant this warp may touch
    const int row_in_tile = quad * 32 + lane;    // TMEM lane == query row inside this CTA's tile
    int* keys_row = reinterpret_cast<int*>(smem + L::topv) + row_in_tile;   // [kSlots][BM] packed keys
    const float window_scaled = 2.0f * filter_eps(q_err, p_err, k_blocks * BK) * kDotScale;
    uint32_t tile_n = 0;
    int it = 0, sq = 0;
    uint32_t sq_phase = 0;
    for (;;) {
      int u;
      if (dyn) {
        mbar_wait(sbase + L::sched_full_bar + sq * 8, sq_phase);
        u = *reinterpret_cast<volatile int*>(smem + L::sched_unit + sq * 4);
        __syncwarp();
        if (lane == 0) mbar_arrive(sbase + L::sched_empty_bar + sq * 8);
        if (++sq == kSchedDepth) {
          sq = 0;
          sq_phase ^= 1;
        }
      } else {
        u = worker + (it++) * n_workers;
        if (u >= n_units) u = -1;
      }
      if (u < 0) break;
      const Unit un = decode_unit(u, n_qtiles, n_seg, n_blk, n_ptiles, grp);
      const int seg = un.seg, qt = un.qt, t0 = un.t0, t1 = un.t1;
      const int64_t row = (int64_t)qt * BM + row_in_tile;
      const bool row_ok = row < n_query;
      const int flag_base = (qt * 4 + quad) * n_seg;
      const int64_t slot = row * n_seg + seg;
      RowState st;
      if (un.blk > 0) {
        // The chain's earlier blocks ran as earlier units, possibly on another SM.  Wait until the
        // last of them has published (this warp's counter of finished blocks; in steady state it
        // has, long ago), then pick the rows' state up from global memory.  The wait always ends:
        // a unit only waits for a LOWER unit index, units are claimed in index order by CTAs that
        // are already running, and the lowest unfinished unit therefore never waits.
        if (lane == 0) {
          while (*reinterpret_cast<volatile int*>(seg_flag + flag_base + seg) < un.blk) __nanosleep(200);
        }
        __syncwarp();
        __threadfence();
      }
      if (un.blk > 0 && row_ok) {
        int kmin = INT_MAX;
        for (int j = 0; j < kSlots; ++j) {
          const int key = j < k ? pack_key(__ldcg(seg_top + slot * k + j) * kDotScale, j) : INT_MAX;
          keys_row[j * BM] = key;
          kmin = min(kmin, key);
        }
        st.kth = key_value(kmin);
        st.kpos = kmin & 31;
        st.cnt = __ldcg(log_cnt + slot);
      } else {
        for (int j = 0; j < kSlots; ++j) keys_row[j * BM] = j < k ? pack_key(kEmpty, j) : INT_MAX;
        st.kth = kEmpty;
        st.kpos = 0;
        st.cnt = 0;
      }
      // Warm start: the k-th best similarity any OTHER segment of this row has published is a lower
      // bound of the row's global k-th best, so nothing more than 2*eps below it can be a true
      // neighbour (counter written after the values, both fenced).
      float known = kEmpty;
      if (row_ok) {
        for (int s2 = 0; s2 < n_seg; ++s2) {
          if (s2 == seg) continue;
          if (*reinterpret_cast<volatile int*>(seg_flag + flag_base + s2)) {
            __threadfence();
            known = fmaxf(known, __ldcg(seg_kth + row * n_seg + s2));
          }
        }
      }
      const float warm_lo = row_ok ? known * kDotScale - window_scaled : INFINITY;
      st.tau_lo = fmaxf(warm_lo, st.kth - window_scaled);
      int m_lo = 0, m_hi = 0;   // masked column range of this row (empty unless MASKED)
      if constexpr (MASKED) {
        if (row_ok) {
          m_lo = (int)mask_lo[row];
          m_hi = (int)mask_hi[row];
        }
      }
      float* lv = log_val + (row_ok ? slot * cap : 0);
      int* li = log_idx + (row_ok ? slot * cap : 0);
      for (int pt = t0; pt < t1; ++pt, ++tile_n) {
        const uint32_t buf = tile_n & 1;
        const uint32_t buf_phase = (tile_n >> 1) & 1;
        mbar_wait_backoff(sbase + L::tmem_full_bar + buf * 8, buf_phase, spin_ns >> 16);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * BN;
        const int col0 = pt * BN;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
          if constexpr (MASKED) {
            const int cb = col0 + c * 32;
            if (cb < m_hi && cb + 32 > m_lo) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (cb + j >= m_lo && cb + j < m_hi) r[j] = 0u;   // similarity 0 <=> distance 1
            }
          }
          float mx = __uint_as_float(r[0]);
#pragma unroll
          for (int j = 1; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
          if (__any_sync(0xffffffffu, mx > st.tau_lo)) {
            // Rare path, lane-parallel: every lane walks ITS OWN passing columns (a divergent loop:
            // the trip count is the largest per-lane count, not the size of the warp's union), picking
            // the value out of its registers with a select tree so the array is never indexed
            // dynamically.  The threshold only rises, so later columns of the chunk are re-checked.
            const int cbase = col0 + c * 32;
            uint32_t mask = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) mask |= (__uint_as_float(r[j]) > st.tau_lo ? 1u : 0u) << j;
            while (mask) {
              const int j = __ffs(mask) - 1;
              mask &= mask - 1;
              const float v = __uint_as_float(select32(r, j));
              if (v > st.tau_lo && (int64_t)(cbase + j) < n_pool && st.cnt <= cap) {
                st = filter_insert(st, v, cbase + j, keys_row, k, lv, li, cap, window_scaled);
                st.tau_lo = fmaxf(st.tau_lo, warm_lo);
              }
            }
            __syncwarp();
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(sbase + L::tmem_empty_bar + buf * 8);
        }
      }
      if (row_ok) {
        log_cnt[slot] = st.cnt;
        for (int j = 0; j < k; ++j) seg_top[slot * k + j] = key_value(keys_row[j * BM]) * kDotUnscale;
        seg_kth[slot] = st.kth * kDotUnscale;
      }
      __threadfence();
      __syncwarp();
      if (lane == 0) *reinterpret_cast<volatile int*>(seg_flag + flag_base + seg) = un.blk + 1;
    }
  }

  // every warp is done with TMEM before its owner frees it
  __syncwarp();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ----------------------------------------------------------------------------- host
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_half_map(CUtensorMap* map, const void* base, int64_t rows, int dim_pad, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  KNN_CHECK_ARG(fn != nullptr, -10, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {(cuuint64_t)dim_pad, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)dim_pad * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  KNN_CHECK_ARG(r == CUDA_SUCCESS, -11, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

int num_sms() {
  static PerDevice cached;   // SM count of each device (0 = not asked yet)
  std::atomic<int>* slot = cached.slot();
  if (slot && slot->load(std::memory_order_relaxed) > 0) return slot->load(std::memory_order_relaxed);
  int dev = 0, v = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0)
    v = 148;
  if (slot) slot->store(v, std::memory_order_relaxed);
  return v;
}

}  // namespace

FilterPlan plan_filter(int64_t n_query, int64_t n_pool, int k) {
  FilterPlan pl;
  pl.ctas = 1;   // CTAs per tcgen05.mma (kept in the plan record for the tests / tools that print it)
  pl.n_qtiles = (int)ceil_div64(n_query, BM);
  pl.n_ptiles = (int)ceil_div64(n_pool, BN);
  const int workers = num_sms();
  // Blocks of <= blk pool tiles (default 96 tiles = 24576 rows = 48 MB of fp16 operand at 1024 dims).
  // Measured on B200 (tools/block_ab.py, 100k x 2M..10M): speed is flat from 64 tiles up; DRAM
  // traffic per launch has a shallow minimum around 96-192 tiles (larger blocks: fewer query-tile
  // re-reads, but the block itself stops surviving in L2 between rounds).
  const int blk = opt_block_tiles() > 0 ? opt_block_tiles() : 96;
  // Number of pool segments (independent chains per query tile): only needed to find work for all
  // CTAs when there are few query tiles.  Cost model: units run `par` at a time (the blocks of a
  // chain are serial, so no more than one unit per chain at once); every segment restarts the
  // running threshold (more log traffic), hence the small per-segment penalty.
  int best_s = 1, best_nb = 1;
  double best_cost = 1e300;
  const int max_s = pl.n_ptiles < 16 ? pl.n_ptiles : 16;
  for (int s = 1; s <= max_s; ++s) {
    const int64_t seg_tiles = ceil_div64(pl.n_ptiles, s);
    const int64_t nb = ceil_div64(seg_tiles, blk);
    const int64_t chains = (int64_t)pl.n_qtiles * s;
    const int64_t par = chains < workers ? chains : workers;
    const int64_t units = chains * nb;
    const double tiles = (double)ceil_div64(seg_tiles, nb) + 0.25;  // + state hand-over per unit
    // dynamic claiming: no whole-wave quantisation, about half a unit of tail
    double rounds = units > par ? (double)units / (double)par + 0.5 : 1.0;
    double cost = rounds * tiles * (1.0 + 0.004 * s);
    // with fewer than two chains per CTA the next block of a chain is usually claimed before its
    // predecessor has finished, and the round proceeds at the pace of the slowest SM
    if (nb > 1 && chains < 2 * (int64_t)workers) cost *= 1.10;
    if (cost < best_cost) {
      best_cost = cost;
      best_s = s;
      best_nb = (int)nb;
    }
  }
  pl.n_seg = best_s;
  pl.n_blk = best_nb;
  // chains per group of the two-level unit order (see decode_unit).  Default: ONE group (flat block-major
  // order).  Measured on 100k x 2.5M (profiles/r2_filter_traffic_ab.txt): groups of 148 / 296 / 444 / 592
  // query tiles with 96 / 48 / 32-tile blocks move 58-74 GB per launch against 57-60 GB flat, at the same or
  // lower speed — L2 does not retain a 38-76 MB query group next to the streaming pool block, so the
  // re-reads the grouping was meant to remove stay, and a block is additionally read once per group.
  {
    const int n_chains = pl.n_qtiles * best_s;
    int grp = opt_query_group() > 0 ? opt_query_group() : n_chains;
    if (grp > n_chains) grp = n_chains;
    if (grp < 1) grp = 1;
    pl.grp = grp;
  }
  pl.n_units = pl.n_qtiles * pl.n_seg * pl.n_blk;
  pl.grid = pl.n_units < workers ? pl.n_units : workers;
  // Candidate-log slots per (row, segment).  Sparse data (i.i.d. rows) logs ~70 entries per row at k = 4;
  // WavLM-like rows (a large shared mean: every pool row within ~0.03 cosine of every other) hold
  // ~800-1200 candidates inside the 2*eps window of the k-th best at 1M-10M pool rows, and a row that
  // overflows its log costs a brute-force pass, so the log is sized for the dense case at every k.
  pl.cap = opt_log_cap() > 0 ? opt_log_cap() : 2048;
  return pl;
}

size_t filter_flag_count(const FilterPlan& pl) { return (size_t)pl.n_qtiles * 4 * pl.n_seg; }

template <bool MASKED>
static int launch_variant(const CUtensorMap& map_q, const CUtensorMap& map_p, int64_t n_query, int64_t n_pool,
                          int k_blocks, int k, const FilterPlan& pl, float* log_val, int* log_idx, int* log_cnt,
                          float* seg_top, float* seg_kth, int* seg_flag, const int64_t* mask_lo,
                          const int64_t* mask_hi, const float* q_err, const float* p_err, int* unit_counter,
                          cudaStream_t stream) {
  static PerDevice smem_attr;   // one per template instance
  KNN_SMEM_ATTR(smem_attr, knn_filter_kernel<MASKED>, Cfg::SMEM_BYTES);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(pl.grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  count_launch();
  // a_format/b_format (bits 7-9, 10-12): 0 = fp16, 1 = bf16
  const uint32_t idesc = InstrDesc::value | (opt_bf16() ? ((1u << 7) | (1u << 10)) : 0u);
  KNN_CUDA(cudaLaunchKernelEx(&cfg, knn_filter_kernel<MASKED>, map_q, map_p, n_query, n_pool, k_blocks, k,
                              pl.n_qtiles, pl.n_ptiles, pl.n_seg, pl.n_blk, pl.cap, log_val, log_idx, log_cnt, seg_top, seg_kth,
                              seg_flag, idesc, (uint32_t)opt_spin_ns() | ((uint32_t)opt_epi_sleep_ns() << 16), mask_lo, mask_hi, q_err, p_err,
                              unit_counter, opt_filter_flags(), pl.grp));
  return 0;
}

int launch_knn_filter(const void* qh, int64_t n_query, const void* ph, int64_t n_pool, int dim_pad, int k,
                      const FilterPlan& pl, float* log_val, int* log_idx, int* log_cnt, float* seg_top,
                      float* seg_kth, int* seg_flag, const int64_t* mask_lo, const int64_t* mask_hi,
                      const float* q_err, const float* p_err, int* unit_counter, cudaStream_t stream) {
  KNN_CHECK_ARG((mask_lo == nullptr) == (mask_hi == nullptr), -3, "mask_lo and mask_hi must be given together");
  KNN_CHECK_ARG(unit_counter != nullptr, -3, "unit counter missing");
  KNN_CHECK_ARG(dim_pad % BK == 0 && dim_pad > 0, -3, "dim_pad %d must be a positive multiple of %d", dim_pad, BK);
  KNN_CHECK_ARG(k >= 1 && k <= kMaxK, -3, "k=%d outside [1,%d]", k, kMaxK);
  KNN_CHECK_ARG(n_pool < (int64_t)1 << 31, -3, "pool shard of %lld rows exceeds int32 column indices", (long long)n_pool);
  CUtensorMap map_q, map_p;
  int rc = make_half_map(&map_q, qh, n_query, dim_pad, BM);
  if (rc) return rc;
  rc = make_half_map(&map_p, ph, n_pool, dim_pad, BN);
  if (rc) return rc;
  if (mask_lo)
    return launch_variant<true>(map_q, map_p, n_query, n_pool, dim_pad / BK, k, pl, log_val, log_idx, log_cnt, seg_top,
                                seg_kth, seg_flag, mask_lo, mask_hi, q_err, p_err, unit_counter, stream);
  return launch_variant<false>(map_q, map_p, n_query, n_pool, dim_pad / BK, k, pl, log_val, log_idx, log_cnt, seg_top,
                               seg_kth, seg_flag, mask_lo, mask_hi, q_err, p_err, unit_counter, stream);
}

}  // namespace knnsvc
