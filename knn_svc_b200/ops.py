"""Tensor-level ops: torch tensors in, C-ABI call on the current CUDA stream,
torch tensors out.  PyTorch is used for device memory and streams only; all
arithmetic happens in libknnsvc_b200.so.  Every op requires CUDA tensors and
raises otherwise — there is no CPU path.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import _lib

MAX_K = 32
WORKSPACE_BUDGET = 16 << 30    # bytes of search scratch above which the query rows are chunked
_MIN_CHUNK = 128 * 148         # one query tile per SM
_HALF_ALIGN = 64   # the tcgen05 filter consumes the feature dimension in 64-element slabs


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _dev(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"{what} must be a CUDA tensor: knn_svc_b200 has no CPU fallback")


def _f32c(t: torch.Tensor) -> torch.Tensor:
    """fp32 contiguous view/copy.  The kernels take fp32 rows.  float64 inputs are the reference's
    real inference path (a float64 one-hot layer weighting promotes fp32 WavLM features
    to float64 CONTAINERS of fp32-representable values), so the narrowing is lossless there — and
    that is checked: a float64 tensor that does not survive the round trip through fp32 raises
    instead of being silently rounded (the reference would have computed on the fp64 values).
    fp16 / bf16 inputs widen losslessly."""
    if t.dtype == torch.float64:
        t32 = t.to(torch.float32)
        if t.numel() and not torch.equal(t32.to(torch.float64), t):
            raise ValueError("float64 input holds values that are not representable in float32; "
                             "the CUDA path computes on fp32 rows (SURVEY D8) and will not round them silently")
        t = t32
    elif t.dtype != torch.float32:
        t = t.to(torch.float32)
    return t.contiguous()


def _i64c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.int64:
        t = t.to(torch.int64)
    return t.contiguous()


def _ptr(t):
    return None if t is None else t.data_ptr()


@dataclass
class PreparedRows:
    """A [n, dim] fp32 row set with its norms and the fp16 tensor-core operand."""
    rows: torch.Tensor      # [n, dim] fp32
    half: torch.Tensor      # [n, dim_pad] fp16 = fp16(rows * 1024/|row|)
    norms: torch.Tensor     # [n] fp64 (fp64-accumulated |row|)
    err: torch.Tensor = None  # [1] fp32: max over rows of |half/1024 - row/|row||_2 (measured rounding error)

    @property
    def n(self) -> int:
        return self.rows.shape[0]

    @property
    def dim(self) -> int:
        return self.rows.shape[1]

    @property
    def dim_pad(self) -> int:
        return self.half.shape[1]


def prepare_rows(x: torch.Tensor, check: bool = True) -> PreparedRows:
    """Row norms + unit-normalised fp16 copy (replaces torch.norm at
    lib_ongaku_test.py:150-151).  A zero-norm or non-finite row is an error: the
    reference produces NaN distances there and exits (lib_ongaku_test.py:166-169)."""
    _dev(x, "rows")
    if x.dim() != 2:
        raise ValueError(f"expected [n, dim] rows, got shape {tuple(x.shape)}")
    x = _f32c(x)
    n, dim = x.shape
    dim_pad = (dim + _HALF_ALIGN - 1) // _HALF_ALIGN * _HALF_ALIGN
    half = torch.empty((n, dim_pad), dtype=torch.float16, device=x.device)
    norms = torch.empty((n,), dtype=torch.float64, device=x.device)
    scal = torch.zeros((2,), dtype=torch.int32, device=x.device)     # [0] bad-row counter, [1] max error (float bits)
    bad, err = scal[0:1], scal[1:2].view(torch.float32)
    # With `check` the bad-row counter lives in PINNED HOST memory (device-accessible at its host address): the
    # kernel touches it only when it finds a bad row, and reading it costs a stream synchronisation but no
    # device-to-host copy — which would wait on the copy engine behind any large download another stream has
    # queued (a previous batch's features).
    flag = _pinned_flag() if check else None
    if flag is not None:
        flag[0] = 0
    lib = _lib.load()
    if n > 0:
        with torch.cuda.device(x.device):
            _lib.check(lib.knnsvc_prepare_rows(x.data_ptr(), n, dim, dim, half.data_ptr(), dim_pad, norms.data_ptr(),
                                               (flag if flag is not None else bad).data_ptr(), err.data_ptr(),
                                               _stream()), "prepare_rows")
            if flag is not None:
                torch.cuda.current_stream(x.device).synchronize()
                n_bad = int(flag[0])
                if n_bad != 0:
                    raise ValueError(f"{n_bad} zero-norm or non-finite feature rows: cosine distance undefined "
                                     "(the reference exits with 'containing nan')")
    return PreparedRows(x, half, norms, err)


_host_local = None


def _thread_state():
    global _host_local
    import threading
    if _host_local is None:
        _host_local = threading.local()
    return _host_local


def _pinned_flag() -> torch.Tensor:
    """one pinned int32 per host thread"""
    st = _thread_state()
    if getattr(st, "flag", None) is None:
        st.flag = torch.zeros((1,), dtype=torch.int32).pin_memory()
    return st.flag


def to_host_small(t: torch.Tensor) -> torch.Tensor:
    """A small device tensor -> a fresh host tensor WITHOUT the copy engine: a kernel stores it into a pinned staging
    buffer (grow-only, one per host thread), the stream is synchronised, the staging area is copied out.  A blocking
    `.cpu()` would queue behind whatever another stream is downloading (knnsvc_store_to_host)."""
    _dev(t, "tensor")
    t = t.contiguous()
    nbytes = t.numel() * t.element_size()
    if nbytes == 0 or nbytes % 4 != 0 or t.data_ptr() % 4 != 0:      # (the kernel moves 32-bit words)
        return t.cpu()
    st = _thread_state()
    if getattr(st, "staging", None) is None or st.staging.numel() < nbytes:
        st.staging = torch.empty((max(nbytes, 1 << 20),), dtype=torch.uint8).pin_memory()
    lib = _lib.load()
    with torch.cuda.device(t.device):
        _lib.check(lib.knnsvc_store_to_host(t.data_ptr(), st.staging.data_ptr(), nbytes, _stream()), "store_to_host")
        torch.cuda.current_stream(t.device).synchronize()
    return st.staging[:nbytes].clone().view(t.dtype).reshape(t.shape)


def cosine_dist(q: torch.Tensor, p: torch.Tensor) -> torch.Tensor:
    """Full [T, Np] cosine-distance matrix (fast_cosine_dist's return value)."""
    _dev(q, "source_feats"); _dev(p, "matching_pool")
    q, p = _f32c(q), _f32c(p)
    if q.shape[1] != p.shape[1]:
        raise ValueError("feature dimensions differ")
    out = torch.empty((q.shape[0], p.shape[0]), dtype=torch.float32, device=q.device)
    lib = _lib.load()
    with torch.cuda.device(q.device):
        _lib.check(lib.knnsvc_cosine_dist(q.data_ptr(), q.shape[0], p.data_ptr(), p.shape[0], q.shape[1],
                                          out.data_ptr(), _stream()), "cosine_dist")
    return out


_ws_cache: dict = {}


def _workspace(nbytes: int, device) -> torch.Tensor:
    """Grow-only scratch buffer per (device, CUDA stream, host thread): two searches on different
    streams or from different threads never share candidate logs / counters, and a buffer is only
    ever used on the stream it was allocated on.  (The search itself never allocates; the library's
    only internal allocation is K5's small per-call scratch, from its own stream-ordered pool.)"""
    import threading
    index = device.index if device.index is not None else torch.cuda.current_device()
    key = (index, torch.cuda.current_stream(index).cuda_stream, threading.get_ident())
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty((max(nbytes, 1),), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def release_workspaces():
    """Drop every cached search scratch buffer (they are grow-only otherwise)."""
    _ws_cache.clear()


def knn_search(query: PreparedRows, pool: PreparedRows, k: int, index_offset: int = 0,
               return_stats: bool = False, mask_lo: torch.Tensor | None = None,
               mask_hi: torch.Tensor | None = None, return_dist64: bool = False):
    """k smallest cosine distances per query row, ascending, with int64 pool
    indices — the fused replacement of fast_cosine_dist + topk
    (ddsp_prematch_dataset.py:1196-1206, ddsp_matcher.py:550-554).

    `mask_lo` / `mask_hi` ([T] int64, together): the pool columns [mask_lo[t], mask_hi[t]) of
    query row t get distance exactly 1 — the offline prematch's self-utterance rule
    `dists[:, start_index:end_index] = 1` (ddsp_prematch_dataset.py:1623-1624).
    `return_dist64`: also return the fp64 distances the re-score ranked by (the sharded path
    exchanges and merges those; `dist` is their fp32 rounding).  Return order:
    dist, idx[, dist64][, stats]."""
    if (mask_lo is None) != (mask_hi is None):
        raise ValueError("mask_lo and mask_hi must be given together")
    if not (1 <= k <= MAX_K):
        raise ValueError(f"k={k} outside [1,{MAX_K}]")
    if k > pool.n:
        raise ValueError(f"k={k} exceeds the pool size {pool.n}")
    if query.dim != pool.dim:
        raise ValueError("feature dimensions differ")
    dev = query.rows.device
    T = query.n
    dist = torch.empty((T, k), dtype=torch.float32, device=dev)
    idx = torch.empty((T, k), dtype=torch.int64, device=dev)
    dist64 = torch.empty((T, k), dtype=torch.float64, device=dev) if return_dist64 else None
    stats = torch.zeros((8,), dtype=torch.int32, device=dev)
    lib = _lib.load()
    if mask_lo is not None:
        mask_lo, mask_hi = _i64c(mask_lo.to(dev)), _i64c(mask_hi.to(dev))
        if mask_lo.shape != (T,) or mask_hi.shape != (T,):
            raise ValueError("mask_lo / mask_hi must have one entry per query row")
    if T > 0:
        with torch.cuda.device(dev):
            # The candidate log is sized for the worst case (rows x segments x cap); very large query
            # sets are searched in row chunks so the scratch buffer stays below WORKSPACE_BUDGET.
            chunk = T
            while chunk > _MIN_CHUNK and lib.knnsvc_knn_workspace_bytes(chunk, pool.n, query.dim_pad, k) > WORKSPACE_BUDGET:
                chunk = max(_MIN_CHUNK, (chunk // 2 + _MIN_CHUNK - 1) // _MIN_CHUNK * _MIN_CHUNK)
            ws = _workspace(lib.knnsvc_knn_workspace_bytes(chunk, pool.n, query.dim_pad, k), dev)
            part_stats = torch.zeros((8,), dtype=torch.int32, device=dev) if chunk < T else stats
            for a in range(0, T, chunk):
                n = min(chunk, T - a)
                _lib.check(lib.knnsvc_knn_search_full(
                    query.rows[a:].data_ptr(), query.half[a:].data_ptr(), query.norms[a:].data_ptr(), n,
                    pool.rows.data_ptr(), pool.half.data_ptr(), pool.norms.data_ptr(), pool.n, query.dim,
                    query.dim_pad, k, index_offset, _ptr(query.err), _ptr(pool.err),
                    None if mask_lo is None else mask_lo[a:].data_ptr(),
                    None if mask_hi is None else mask_hi[a:].data_ptr(),
                    dist[a:].data_ptr(), None if dist64 is None else dist64[a:].data_ptr(), idx[a:].data_ptr(),
                    ws.data_ptr(), ws.numel(), part_stats.data_ptr(), _stream()), "knn_search")
                if chunk < T:
                    stats[:3] += part_stats[:3]          # flagged rows, logged candidates, survivors
                    stats[3:] = part_stats[3:]
    out = (dist, idx) + ((dist64,) if return_dist64 else ()) + ((stats,) if return_stats else ())
    return out


def knn_workspace_layout(n_query: int, n_pool: int, dim_pad: int, k: int) -> dict:
    """Where knn_search keeps its candidate log inside the scratch buffer (diagnostics / tests)."""
    import ctypes
    arr = (ctypes.c_int64 * 8)()
    _lib.check(_lib.load().knnsvc_knn_workspace_layout(n_query, n_pool, dim_pad, k, ctypes.cast(arr, ctypes.c_void_p)),
               "knn_workspace_layout")
    names = ("log_val", "log_idx", "log_cnt", "seg_top", "n_seg", "cap", "total", "ref_val")
    return {n: int(arr[i]) for i, n in enumerate(names)}


def knn_candidate_log(query: PreparedRows, pool: PreparedRows, k: int):
    """Run the search and hand back what the tensor-core filter logged: (values [T*n_seg, cap] fp32 —
    the accumulator's cosine similarities s~ —, columns [T*n_seg, cap] int32, counts [T*n_seg],
    layout dict), plus the search result.  Test / diagnostic aid: the error-window bound
    |s~ - s| <= eps is checked against THESE values, i.e. against what tcgen05.mma left in TMEM."""
    T = query.n
    res = knn_search(query, pool, k, return_stats=True)
    lay = knn_workspace_layout(T, pool.n, query.dim_pad, k)
    ws = _workspace(lay["total"], query.rows.device)
    slots, cap = T * lay["n_seg"], lay["cap"]
    val = ws[lay["log_val"]:lay["log_val"] + slots * cap * 4].view(torch.float32).view(slots, cap).clone()
    col = ws[lay["log_idx"]:lay["log_idx"] + slots * cap * 4].view(torch.int32).view(slots, cap).clone()
    cnt = ws[lay["log_cnt"]:lay["log_cnt"] + slots * 4].view(torch.int32).clone()
    return val, col, cnt, lay, res


def knn_exact(query: PreparedRows, pool: PreparedRows, k: int, index_offset: int = 0):
    """Exact CUDA-core kNN (fp64 accumulation); the decision procedure behind
    knn_search for undecidable rows, exposed for tests."""
    dev = query.rows.device
    T = query.n
    dist = torch.empty((T, k), dtype=torch.float32, device=dev)
    idx = torch.empty((T, k), dtype=torch.int64, device=dev)
    lib = _lib.load()
    if T > 0:
        with torch.cuda.device(dev):
            nbytes = lib.knnsvc_knn_exact_workspace_bytes(T, pool.n, k)
            ws = _workspace(nbytes, dev)
            _lib.check(lib.knnsvc_knn_exact(query.rows.data_ptr(), query.norms.data_ptr(), T, pool.rows.data_ptr(),
                                            pool.norms.data_ptr(), pool.n, query.dim, k, index_offset,
                                            dist.data_ptr(), idx.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
                       "knn_exact")
    return dist, idx


def merge_topk(gathered_dist: torch.Tensor, gathered_idx: torch.Tensor):
    """[R, T, k] per-shard results -> [T, k] merged by (dist, idx) (C1)."""
    _dev(gathered_dist, "gathered_dist")
    R, T, k = gathered_dist.shape
    gd, gi = _f32c(gathered_dist), _i64c(gathered_idx)
    dist = torch.empty((T, k), dtype=torch.float32, device=gd.device)
    idx = torch.empty((T, k), dtype=torch.int64, device=gd.device)
    lib = _lib.load()
    if T == 0:
        return dist, idx
    with torch.cuda.device(gd.device):
        _lib.check(lib.knnsvc_merge_topk(gd.data_ptr(), gi.data_ptr(), R, T, k, dist.data_ptr(), idx.data_ptr(),
                                         _stream()), "merge_topk")
    return dist, idx


def merge_topk64(gathered_dist64: torch.Tensor, gathered_idx: torch.Tensor):
    """[R, T, k] per-shard fp64 distances + global indices -> ([T,k] fp32, [T,k] fp64, [T,k] int64)
    merged by (fp64 dist, idx): bit for bit what one search of the whole pool returns."""
    _dev(gathered_dist64, "gathered_dist64")
    R, T, k = gathered_dist64.shape
    gd, gi = gathered_dist64.to(torch.float64).contiguous(), _i64c(gathered_idx)
    dist = torch.empty((T, k), dtype=torch.float32, device=gd.device)
    dist64 = torch.empty((T, k), dtype=torch.float64, device=gd.device)
    idx = torch.empty((T, k), dtype=torch.int64, device=gd.device)
    if T == 0:
        return dist, dist64, idx
    lib = _lib.load()
    with torch.cuda.device(gd.device):
        _lib.check(lib.knnsvc_merge_topk64(gd.data_ptr(), gi.data_ptr(), R, T, k, dist.data_ptr(), dist64.data_ptr(),
                                           idx.data_ptr(), _stream()), "merge_topk64")
    return dist, dist64, idx


class ShardedRows:
    """Row table of a pool that lives in several blocks — this GPU's shard plus its peers'
    IPC-mapped shards: `ptrs[s]` is a device pointer valid in THIS process for global rows
    [bounds[s], bounds[s+1])."""

    def __init__(self, ptrs, bounds, dim: int, device):
        import ctypes
        assert len(bounds) == len(ptrs) + 1
        self.n, self.dim, self.device = len(ptrs), int(dim), torch.device(device)
        self.ptrs, self.bounds = [int(p) for p in ptrs], [int(b) for b in bounds]
        self._ptr_arr = (ctypes.c_void_p * self.n)(*self.ptrs)
        self._lo_arr = (ctypes.c_int64 * (self.n + 1))(*self.bounds)


def gather_mix_sharded(table: ShardedRows, idx: torch.Tensor, weights: torch.Tensor | None = None) -> torch.Tensor:
    """gather_mix over a sharded pool (GLOBAL indices); bit-identical to gather_mix on the
    concatenated pool.  Rows of other GPUs are read over NVLink through their mapped pointers."""
    import ctypes
    _dev(idx, "indices")
    idx = _i64c(idx)
    T, k = idx.shape
    w = None if weights is None else _f32c(weights.to(idx.device))
    out = torch.empty((T, table.dim), dtype=torch.float32, device=idx.device)
    if T == 0:
        return out
    lib = _lib.load()
    with torch.cuda.device(idx.device):
        _lib.check(lib.knnsvc_gather_mix_sharded(ctypes.cast(table._ptr_arr, ctypes.c_void_p),
                                                 ctypes.cast(table._lo_arr, ctypes.c_void_p), table.n, table.dim,
                                                 idx.data_ptr(), _ptr(w), T, k, out.data_ptr(), _stream()),
                   "gather_mix_sharded")
    return out


def gather_mix(pool: torch.Tensor, idx: torch.Tensor, weights: torch.Tensor | None = None) -> torch.Tensor:
    """out[t] = sum_k w[t,k] * pool[idx[t,k]]  (weights None -> mean)."""
    _dev(pool, "pool"); _dev(idx, "indices")
    pool, idx = _f32c(pool), _i64c(idx)
    T, k = idx.shape
    w = None if weights is None else _f32c(weights.to(pool.device))
    out = torch.empty((T, pool.shape[1]), dtype=torch.float32, device=pool.device)
    lib = _lib.load()
    if T == 0:
        return out
    with torch.cuda.device(pool.device):
        _lib.check(lib.knnsvc_gather_mix(pool.data_ptr(), pool.shape[0], pool.shape[1], idx.data_ptr(), _ptr(w), T, k,
                                         out.data_ptr(), _stream()), "gather_mix")
    return out


def f0_rerank(expected_f0: torch.Tensor, pool_f0: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    _dev(idx, "indices")
    dev = idx.device
    e, f, idx = _f32c(expected_f0.to(dev)), _f32c(pool_f0.to(dev)), _i64c(idx)
    T, k = idx.shape
    if e.shape[0] != T:
        raise ValueError("expected_f0 and indices disagree on the number of frames")
    out = torch.empty_like(idx)
    lib = _lib.load()
    if T == 0:
        return out
    with torch.cuda.device(dev):
        _lib.check(lib.knnsvc_f0_rerank(e.data_ptr(), f.data_ptr(), idx.data_ptr(), T, k, out.data_ptr(), _stream()),
                   "f0_rerank")
    return out


def concat_cost_reselect(idx: torch.Tensor, src: torch.Tensor, pool, shifted_src_f0: torch.Tensor | None = None,
                         pool_f0: torch.Tensor | None = None, concat_weight: float = 0.2, utt_offsets=None) -> torch.Tensor:
    """Greedy concatenation-cost re-selection (K5).  `pool` is the [Np, D] pool tensor or a
    `ShardedRows` table (a pool sharded over GPUs: `idx` then holds GLOBAL rows and `pool_f0` the f0 of
    the whole pool); results are identical."""
    _dev(src, "src_elements")
    dev = src.device
    table = pool if isinstance(pool, ShardedRows) else None
    if table is None:
        _dev(pool, "tgt_elements")
        pool = _f32c(pool)
    idx, src = _i64c(idx.to(dev)), _f32c(src)
    T, k = idx.shape
    if k != 4:
        raise ValueError("knn_with_concat_cost: the reference path keeps 4 candidates per frame")
    if len(src) != T:
        raise ValueError("indices and src_elements disagree on the number of frames")
    sf = None if shifted_src_f0 is None else _f32c(shifted_src_f0.to(dev))
    pf = None if pool_f0 is None else _f32c(pool_f0.to(dev))
    offs = [0, T] if utt_offsets is None else [int(v) for v in utt_offsets]
    import ctypes
    arr = (ctypes.c_int64 * len(offs))(*offs)
    out = torch.empty_like(idx)
    lib = _lib.load()
    if T == 0:
        return out
    with torch.cuda.device(dev):
        if table is None:
            _lib.check(lib.knnsvc_concat_cost_reselect(idx.data_ptr(), src.data_ptr(), pool.data_ptr(), pool.shape[0],
                                                       pool.shape[1], _ptr(sf), _ptr(pf), float(concat_weight),
                                                       ctypes.cast(arr, ctypes.c_void_p), len(offs) - 1, out.data_ptr(),
                                                       _stream()), "concat_cost_reselect")
        else:
            if src.shape[1] != table.dim:
                raise ValueError("feature dimensions differ")
            _lib.check(lib.knnsvc_concat_cost_reselect_sharded(
                idx.data_ptr(), src.data_ptr(), ctypes.cast(table._ptr_arr, ctypes.c_void_p),
                ctypes.cast(table._lo_arr, ctypes.c_void_p), table.n, table.dim, _ptr(sf), _ptr(pf), float(concat_weight),
                ctypes.cast(arr, ctypes.c_void_p), len(offs) - 1, out.data_ptr(), _stream()), "concat_cost_reselect_sharded")
    return out


def weight_fit(idx: torch.Tensor, synth, loss_scale: float, max_iters: int = 100000,
               return_info: bool = False, utt_offsets=None, amp_ratio: torch.Tensor | None = None):
    """Adam(amsgrad) fit of the softmax mixing weights (K6).  With `utt_offsets` the rows of
    `idx` are a concatenation of utterances, each fitted independently (one CTA each, one
    launch); info is then [n_utt, 4].  `amp_ratio` [T,k] scales every candidate row before
    mixing (compute_weight_with_amp, ddsp_prematch_dataset.py:684-803).  `synth` is the [Np, D]
    tensor or a `ShardedRows` table (GLOBAL indices; no amp_ratio)."""
    table = synth if isinstance(synth, ShardedRows) else None
    if table is None:
        _dev(synth, "synth_set")
        synth = _f32c(synth)
        dev = synth.device
    else:
        dev = table.device
    idx = _i64c(idx.to(dev))
    T, k = idx.shape
    amp = None
    if amp_ratio is not None:
        if tuple(amp_ratio.shape) != (T, k):
            raise AssertionError("amp_ratio must have the shape of target_feature_indices")   # reference :687
        if table is not None:
            raise ValueError("amp_ratio is not supported on a sharded pool")
        amp = _f32c(amp_ratio.to(dev))
    offs = [0, T] if utt_offsets is None else [int(v) for v in utt_offsets]
    if offs[0] != 0 or offs[-1] != T:
        raise ValueError("utt_offsets must run from 0 to the number of frames")
    n_utt = len(offs) - 1
    out = torch.empty((T, k), dtype=torch.float32, device=dev)
    info = torch.zeros((n_utt, 4), dtype=torch.float64, device=dev)
    lib = _lib.load()
    if T > 0:
        import ctypes
        arr = (ctypes.c_int64 * len(offs))(*offs)
        with torch.cuda.device(dev):
            nbytes = lib.knnsvc_weight_fit_batched_workspace_bytes(T, k, n_utt)
            ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
            if table is None:
                _lib.check(lib.knnsvc_weight_fit_amp(idx.data_ptr(), synth.data_ptr(), synth.shape[0], synth.shape[1],
                                                     ctypes.cast(arr, ctypes.c_void_p), n_utt, k, float(loss_scale),
                                                     int(max_iters), _ptr(amp), out.data_ptr(), info.data_ptr(),
                                                     ws.data_ptr(), ws.numel(), _stream()), "weight_fit")
            else:
                _lib.check(lib.knnsvc_weight_fit_sharded(idx.data_ptr(), ctypes.cast(table._ptr_arr, ctypes.c_void_p),
                                                         ctypes.cast(table._lo_arr, ctypes.c_void_p), table.n, table.dim,
                                                         ctypes.cast(arr, ctypes.c_void_p), n_utt, k, float(loss_scale),
                                                         int(max_iters), out.data_ptr(), info.data_ptr(), ws.data_ptr(),
                                                         ws.numel(), _stream()), "weight_fit_sharded")
    if utt_offsets is None:
        info = info[0]
    if return_info:
        return out, info
    return out


def harmonic_bank(f0: torch.Tensor, amp: torch.Tensor | None, sample_rate: int = 16000, hop: int = 320) -> torch.Tensor:
    """f0 [B,T], amp [B,T,H] or None -> [B, T*hop] fp32."""
    _dev(f0, "f0")
    f0 = _f32c(f0)
    B, T = f0.shape
    H = 1
    if amp is not None:
        amp = _f32c(amp.to(f0.device))
        if amp.shape[:2] != f0.shape:
            raise ValueError("f0 and amp disagree on [batch, frames]")
        H = amp.shape[2]
    out = torch.empty((B, T * hop), dtype=torch.float32, device=f0.device)
    ws = torch.empty((max(B * T, 1),), dtype=torch.float64, device=f0.device)
    lib = _lib.load()
    if B * T == 0:
        return out
    with torch.cuda.device(f0.device):
        _lib.check(lib.knnsvc_harmonic_bank(f0.data_ptr(), _ptr(amp), B, T, H, int(sample_rate), int(hop),
                                            out.data_ptr(), ws.data_ptr(), _stream()), "harmonic_bank")
    return out


# ----------------------------------------------------------------------------- SURVEY §8(f) ops


def layer_mix(feats: torch.Tensor, weights_a, weights_b=None):
    """`(feats*weights[:, None]).sum(dim=0)` for one or two weight vectors in one pass over the
    [L, T, D] layer stack (ddsp_prematch_dataset.py:349-350).  Weights are host float64."""
    _dev(feats, "feats")
    feats = _f32c(feats)
    L, T, D = feats.shape
    import ctypes
    import numpy as np
    wa = np.ascontiguousarray(np.asarray(weights_a, dtype=np.float64).reshape(-1))
    if wa.shape[0] != L:
        raise ValueError("one weight per layer expected")
    wb = None
    out_a = torch.empty((T, D), dtype=torch.float32, device=feats.device)
    out_b = None
    if weights_b is not None:
        wb = np.ascontiguousarray(np.asarray(weights_b, dtype=np.float64).reshape(-1))
        if wb.shape[0] != L:
            raise ValueError("one weight per layer expected")
        out_b = torch.empty_like(out_a)
    lib = _lib.load()
    if T * D == 0:
        return out_a if weights_b is None else (out_a, out_b)
    with torch.cuda.device(feats.device):
        _lib.check(lib.knnsvc_layer_mix(feats.data_ptr(), L, T, D, wa.ctypes.data_as(ctypes.c_void_p),
                                        None if wb is None else wb.ctypes.data_as(ctypes.c_void_p),
                                        out_a.data_ptr(), _ptr(out_b), _stream()), "layer_mix")
    return out_a if weights_b is None else (out_a, out_b)


def stft_magnitude(audio: torch.Tensor, frames: int | None = None, n_fft: int = 400, hop: int = 320) -> torch.Tensor:
    """Spectrogram(n_fft, hop, center=True, power=1)(x).T[:, :-1][:frames] — reference :326, :361-363."""
    _dev(audio, "audio")
    x = _f32c(audio.reshape(-1))
    n = x.shape[0]
    total = 1 + n // hop
    frames = total if frames is None else int(frames)
    if frames > total:
        raise AssertionError("spectrogram shorter than the feature sequence")       # reference :362
    out = torch.empty((frames, n_fft // 2), dtype=torch.float32, device=x.device)
    lib = _lib.load()
    if frames == 0:
        return out
    with torch.cuda.device(x.device):
        _lib.check(lib.knnsvc_stft_magnitude(x.data_ptr(), n, frames, n_fft, hop, out.data_ptr(), _stream()),
                   "stft_magnitude")
    return out


def harmonic_amplitudes(spec: torch.Tensor, f0: torch.Tensor, n_harm: int = 49, sample_rate: int = 16000) -> torch.Tensor:
    """0.0108 * spectrum magnitude at the first `n_harm` harmonics of f0 — reference :391-404."""
    _dev(spec, "spec")
    spec = _f32c(spec)
    f0 = _f32c(f0.to(spec.device))
    T, S = spec.shape
    if f0.shape != (T,):
        raise ValueError("one f0 value per spectrum row expected")
    if sample_rate / (2 * S) != 40:
        raise AssertionError([sample_rate / (2 * S)])                                # reference :392
    out = torch.empty((T, n_harm), dtype=torch.float32, device=spec.device)
    lib = _lib.load()
    if T == 0:
        return out
    with torch.cuda.device(spec.device):
        _lib.check(lib.knnsvc_harmonic_amplitudes(spec.data_ptr(), f0.data_ptr(), T, S, n_harm, int(sample_rate),
                                                  out.data_ptr(), _stream()), "harmonic_amplitudes")
    return out


def row_l1(x: torch.Tensor) -> torch.Tensor:
    """`x.norm(dim=1, p=1)` (reference :1672)."""
    _dev(x, "rows")
    x = _f32c(x)
    out = torch.empty((x.shape[0],), dtype=torch.float32, device=x.device)
    lib = _lib.load()
    if x.shape[0] == 0:
        return out
    with torch.cuda.device(x.device):
        _lib.check(lib.knnsvc_row_l1(x.data_ptr(), x.shape[0], x.shape[1], out.data_ptr(), _stream()), "row_l1")
    return out


def amp_ratio(l1_query: torch.Tensor, l1_pool: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """amp_ratio[t,k] = l1_query[t] / (l1_pool[idx[t,k]] + 1e-5) — reference :1672-1675."""
    _dev(l1_pool, "l1_pool")
    dev = l1_pool.device
    q, p, idx = _f32c(l1_query.to(dev)), _f32c(l1_pool), _i64c(idx.to(dev))
    T, k = idx.shape
    if q.shape != (T,):
        raise ValueError("one query norm per index row expected")
    out = torch.empty((T, k), dtype=torch.float32, device=dev)
    lib = _lib.load()
    if T == 0:
        return out
    with torch.cuda.device(dev):
        _lib.check(lib.knnsvc_amp_ratio(q.data_ptr(), p.data_ptr(), idx.data_ptr(), T, k, p.shape[0], out.data_ptr(),
                                        _stream()), "amp_ratio")
    return out
