"""Pool-sharded matching across the GPUs of one box.

One process per GPU.  Rank r owns pool rows [r*Np/R, (r+1)*Np/R); queries are replicated.

  1. every rank runs the fused distance/top-k on its shard (tcgen05 filter + exact re-score);
  2. ONE all-gather (NCCL over NVLink) exchanges k x (fp64 distance, int64 global index) per query
     frame and every rank merges them, ranking on the SAME fp64 distances the single-GPU re-score
     ranks on, ties to the lower global index — so the merged (distances, indices) are bit for bit
     what one search of the whole pool returns, whatever the shard count;
  3. the matched features `synth_set[idx].mean(1)` (ddsp_matcher.py:578,
     ddsp_prematch_dataset.py:1348,1364) are produced by the rank that OWNS the query rows
     (contiguous slices of the batch): its gather kernel reads the k selected pool rows straight
     from whichever GPU holds them, through CUDA-IPC-mapped peer pointers (P2P loads over
     NVLink inside the kernel — no [T, D] all-reduce, and the same arithmetic in the same order as
     the single-GPU gather, so the features are bit-identical too).  `exchange="reduce_scatter"`
     is the NCCL-only alternative (per-shard partial sums + reduce-scatter; equal to ~1 ulp).

The reference has no counterpart: its matcher is single-device
(ddsp_prematch_dataset.py:1196-1206).  The gloo CPU tests cover the host logic (bounds, payload
packing, merge rule, query ownership); everything that computes needs the CUDA library.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import torch
import torch.distributed as dist


def shard_bounds(n_pool: int, world_size: int, rank: int):
    """Contiguous, balanced split by pool frame."""
    lo = (n_pool * rank) // world_size
    hi = (n_pool * (rank + 1)) // world_size
    return lo, hi


def query_slice(n_query: int, world_size: int, rank: int):
    """Query rows whose matched features rank `rank` produces: equal chunks of ceil(T/R) rows
    (the last ones may be short or empty), so slices can be all-gathered without a size exchange."""
    chunk = (n_query + world_size - 1) // world_size
    return min(n_query, rank * chunk), min(n_query, (rank + 1) * chunk)


def _pack(local_dist: torch.Tensor, local_idx: torch.Tensor) -> torch.Tensor:
    """One int64 payload [T, 2k] = (distance bit patterns | indices): one collective for both."""
    T, k = local_dist.shape
    payload = torch.empty((T, 2 * k), dtype=torch.int64, device=local_dist.device)
    if local_dist.dtype == torch.float64:
        payload[:, :k] = local_dist.contiguous().view(torch.int64)
    else:
        payload[:, :k] = local_dist.contiguous().view(torch.int32).to(torch.int64)
    payload[:, k:] = local_idx
    return payload


def all_gather_topk(local_dist: torch.Tensor, local_idx: torch.Tensor, group=None):
    """[T,k] per rank -> ([R,T,k], [R,T,k]) on every rank with ONE all-gather.  Distances may be
    fp32 or fp64 (the sharded search exchanges fp64); the dtype is preserved."""
    world = dist.get_world_size(group)
    T, k = local_dist.shape
    out = torch.empty((world * T, 2 * k), dtype=torch.int64, device=local_dist.device)
    dist.all_gather_into_tensor(out, _pack(local_dist, local_idx), group=group)
    out = out.view(world, T, 2 * k)
    if local_dist.dtype == torch.float64:
        gd = out[:, :, :k].contiguous().view(torch.float64)
    else:
        gd = out[:, :, :k].to(torch.int32).view(torch.float32)
    return gd.contiguous(), out[:, :, k:].contiguous()


def merge_topk_host(gd: torch.Tensor, gi: torch.Tensor):
    """Merge rule in plain torch, (dist, idx) lexicographic — used by the gloo CPU
    test of the exchange logic; GPUs use ops.merge_topk / ops.merge_topk64 (the CUDA kernel)."""
    R, T, k = gd.shape
    d = gd.permute(1, 0, 2).reshape(T, R * k)
    i = gi.permute(1, 0, 2).reshape(T, R * k)
    valid = i >= 0
    d = torch.where(valid, d, torch.full_like(d, float("inf")))
    order = torch.argsort(i, dim=1, stable=True)          # secondary key first
    d, i = torch.gather(d, 1, order), torch.gather(i, 1, order)
    order = torch.argsort(d, dim=1, stable=True)          # then primary key, stably
    return torch.gather(d, 1, order)[:, :k].contiguous(), torch.gather(i, 1, order)[:, :k].contiguous()


@dataclass
class ShardedMatch:
    dist: torch.Tensor        # [T, k] fp32, merged, on every rank
    idx: torch.Tensor         # [T, k] int64 GLOBAL pool indices, merged, on every rank
    feats: torch.Tensor       # matched features: [hi-lo, D] (gather="slice") or [T, D] (gather="all")
    rows: tuple               # (lo, hi): the query rows this rank produced the features of
    dist64: torch.Tensor = None


class ShardedPool:
    """This rank's slice of the target pool, prepared for the tensor-core filter, plus the peer
    table that lets this rank read matched rows from the other ranks' slices."""

    def __init__(self, shard_rows: torch.Tensor, global_offset: int, group=None, exchange: str = "p2p",
                 synth_rows: torch.Tensor | None = None, distributed: bool | None = None):
        """`distributed=False` builds a purely local pool even inside an initialised process group
        (independent replicas); None = distributed iff the group has more than one rank."""
        from . import ops
        if exchange not in ("p2p", "reduce_scatter"):
            raise ValueError("exchange must be 'p2p' or 'reduce_scatter'")
        self.prepared = ops.prepare_rows(shard_rows)
        # the rows features are mixed from (the reference's synth_set; the matching rows unless given)
        self.synth = self.prepared.rows if synth_rows is None else ops._f32c(synth_rows.to(self.prepared.rows.device))
        if self.synth.shape[0] != self.prepared.n:
            raise ValueError("synth_rows must have one row per matching row")
        self.offset = int(global_offset)
        self.group = group
        self.exchange = exchange
        self.device = self.prepared.rows.device
        self.distributed = dist.is_initialized() and dist.get_world_size(group) > 1 if distributed is None \
            else bool(distributed)
        self.world = dist.get_world_size(group) if self.distributed else 1
        self.rank = dist.get_rank(group) if self.distributed else 0
        self._opened = []
        self.table = None
        self.bounds = [self.offset, self.offset + self.prepared.n]
        if self.distributed:
            self._exchange_layout()

    # ------------------------------------------------------------------ construction
    def _exchange_layout(self):
        """Shard bounds of every rank and (exchange == 'p2p') the peer pointers of their rows."""
        from . import _lib, ops
        lib = _lib.load()
        handle = (ctypes.c_ubyte * 64)()
        off = ctypes.c_int64(0)
        if self.exchange == "p2p" and self.synth.numel():
            with torch.cuda.device(self.device):
                _lib.check(lib.knnsvc_ipc_export(self.synth.data_ptr(), ctypes.cast(handle, ctypes.c_void_p),
                                                 ctypes.cast(ctypes.pointer(off), ctypes.c_void_p)), "ipc_export")
        mine = (self.offset, self.prepared.n, bytes(handle), int(off.value), int(self.synth.shape[1]))
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=self.group)
        order = sorted(range(self.world), key=lambda r: everyone[r][0])
        bounds, ptrs = [], []
        for r in order:
            o, n, h, byte_off, dim = everyone[r]
            if bounds and bounds[-1] != o:
                raise ValueError("pool shards must tile the global row range without gaps or overlaps")
            if not bounds:
                bounds.append(o)
            bounds.append(o + n)
            if dim != self.synth.shape[1]:
                raise ValueError("shards disagree on the feature dimension")
            if self.exchange != "p2p":
                continue
            if r == self.rank:
                ptrs.append(self.synth.data_ptr())
            elif n == 0:
                ptrs.append(self.synth.data_ptr())          # never dereferenced: the shard holds no row
            else:
                base = ctypes.c_void_p(0)
                buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
                with torch.cuda.device(self.device):
                    _lib.check(lib.knnsvc_ipc_open(ctypes.cast(buf, ctypes.c_void_p),
                                                   ctypes.cast(ctypes.pointer(base), ctypes.c_void_p)), "ipc_open")
                self._opened.append(base.value)
                ptrs.append(base.value + byte_off)
        self.bounds = bounds
        self.shard_order = order
        if self.exchange == "p2p":
            self.table = ops.ShardedRows(ptrs, bounds, self.synth.shape[1], self.device)
        dist.barrier(group=self.group)      # every peer has mapped every shard before anyone reads

    def close(self):
        """Unmap the peers' shards (collective: nobody may still be reading this rank's rows)."""
        from . import _lib
        if self.distributed and dist.is_initialized():
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)
        lib = _lib.load()
        for base in self._opened:
            lib.knnsvc_ipc_close(ctypes.c_void_p(base))
        self._opened = []
        self.table = None

    def reprepare(self, check: bool = False):
        """Recompute this shard's norms and tensor-core operand from its fp32 rows (what building a
        matching set costs per shard; the rows, and therefore the peers' mappings, stay in place)."""
        from . import ops
        self.prepared = ops.prepare_rows(self.prepared.rows, check=check)

    def upload_query(self, query_host: torch.Tensor) -> torch.Tensor:
        """A query batch in HOST memory -> the replicated [T, D] device tensor every rank searches
        with.  The batch crosses PCIe ONCE in total: rank r uploads its query slice and the ranks
        replicate it over NVLink with one all-gather."""
        T, dim = query_host.shape
        if not self.distributed:
            return query_host.to(self.device, non_blocking=True)
        chunk = (T + self.world - 1) // self.world
        lo, hi = query_slice(T, self.world, self.rank)
        full = torch.empty((self.world * chunk, dim), dtype=query_host.dtype, device=self.device)
        part = torch.zeros((chunk, dim), dtype=query_host.dtype, device=self.device)
        part[:hi - lo].copy_(query_host[lo:hi], non_blocking=True)
        dist.all_gather_into_tensor(full, part, group=self.group)
        return full[:T]

    @property
    def n_total(self) -> int:
        return self.bounds[-1] - self.bounds[0]

    # ------------------------------------------------------------------ search
    def knn(self, query_prepared, k: int, return_dist64: bool = False):
        """Merged k nearest pool rows of every query row, identical on every rank:
        (dist [T,k] fp32, idx [T,k] int64 global)."""
        from . import ops
        if not self.distributed:
            out = ops.knn_search(query_prepared, self.prepared, k, index_offset=self.offset,
                                 return_dist64=return_dist64)
            return out
        if k > self.prepared.n:
            raise ValueError(f"k={k} exceeds this rank's shard of {self.prepared.n} rows")
        _, i, d64 = ops.knn_search(query_prepared, self.prepared, k, index_offset=self.offset, return_dist64=True)
        gd, gi = all_gather_topk(d64, i, self.group)
        d, d64, i = ops.merge_topk64(gd, gi)
        return (d, i, d64) if return_dist64 else (d, i)

    # ------------------------------------------------------------------ the matcher: search + gather-mean
    def row_table(self):
        """The pool as a row table valid on this GPU (peers' shards through their mapped pointers)."""
        from . import ops
        if self.table is not None:
            return self.table
        if self.distributed:
            raise RuntimeError("the peer-memory row table needs exchange='p2p'")
        # one local block; its rows are addressed 0 .. n-1 (match_post_opt removes a non-zero global offset)
        return ops.ShardedRows([self.synth.data_ptr()], [0, self.prepared.n], self.synth.shape[1], self.device)

    def match_post_opt(self, query, concat_weight: float, gather: str = "all", check: bool = True) -> ShardedMatch:
        """`match` followed by the concatenation-smoothness stage on the plain cosine top-4
        (ddsp_prematch_dataset.py:1246,1295,1357-1358): greedy re-selection (skipped when
        `concat_weight` is -1), fitted mixing weights, weighted mix.  `previous selection + 1` and the
        fit's `idx +- 1` rows cross shard boundaries freely: both kernels address the pool through
        the peer row table.  The recurrence is one serial chain per utterance, so every rank runs
        it (deterministic: same bits everywhere); the final gather is split by query slice as in `match`.
        The matching rows and the mixed rows must be the same set (synth_rows=None)."""
        from . import ops
        if self.synth.data_ptr() != self.prepared.rows.data_ptr():
            raise ValueError("post_opt on a sharded pool needs synth_rows to be the matching rows")
        qp = query if isinstance(query, ops.PreparedRows) else \
            ops.prepare_rows(query.to(self.device) if query.is_cuda else self.upload_query(query), check=check)
        d, i, d64 = self.knn(qp, 4, return_dist64=True)
        table = self.row_table()
        local_base = 0 if self.distributed or self.offset == 0 else self.offset
        idx = i - local_base
        if concat_weight != -1:
            idx = ops.concat_cost_reselect(idx, qp.rows, table, concat_weight=concat_weight)
        w = ops.weight_fit(idx, table, 0.1)
        T = qp.n
        lo, hi = query_slice(T, self.world, self.rank) if self.distributed else (0, T)
        part = ops.gather_mix_sharded(table, idx[lo:hi], w[lo:hi])
        idx = idx + local_base
        if gather == "slice" or not self.distributed:
            return ShardedMatch(d, idx, part, (lo, hi), d64)
        chunk = (T + self.world - 1) // self.world
        dim = self.synth.shape[1]
        padded = torch.zeros((chunk, dim), dtype=torch.float32, device=self.device)
        padded[:hi - lo] = part
        allf = torch.empty((self.world * chunk, dim), dtype=torch.float32, device=self.device)
        dist.all_gather_into_tensor(allf, padded, group=self.group)
        return ShardedMatch(d, idx, allf[:T], (0, T), d64)

    def match(self, query, k: int = 4, gather: str = "slice", weights: torch.Tensor | None = None,
              check: bool = True, mix_k: int | None = None) -> ShardedMatch:
        """kNN regression of the query rows onto the sharded pool: the merged top-k and the mean
        (or `weights`-mix) of the k matched `synth` rows — `synth_set[best.indices].mean(dim=1)`,
        ddsp_matcher.py:550-578.  `gather="slice"`: this rank returns the features of ITS query rows
        (`rows`); `gather="all"`: the slices are all-gathered and every rank returns all T rows.
        `mix_k`: mix only the first `mix_k` of the k neighbours — the live path searches k=32 and
        mixes `nearest_nbrs[:, :4]` (ddsp_prematch_dataset.py:1203,1246).  A query in host memory is
        uploaded with `upload_query`."""
        from . import ops
        if gather not in ("slice", "all"):
            raise ValueError("gather must be 'slice' or 'all'")
        if isinstance(query, ops.PreparedRows):
            qp = query
        else:
            qp = ops.prepare_rows(query.to(self.device) if query.is_cuda else self.upload_query(query), check=check)
        T = qp.n
        d, i_all, d64 = self.knn(qp, k, return_dist64=True)
        i = i_all if mix_k is None or mix_k >= k else i_all[:, :mix_k].contiguous()
        k_mix = i.shape[1]
        if not self.distributed:
            return ShardedMatch(d, i_all, ops.gather_mix(self.synth, i - self.offset, weights), (0, T), d64)
        lo, hi = query_slice(T, self.world, self.rank)
        chunk = (T + self.world - 1) // self.world
        dim = self.synth.shape[1]
        if self.exchange == "p2p":
            w = None if weights is None else weights[lo:hi]
            part = ops.gather_mix_sharded(self.table, i[lo:hi], w)
        else:
            # partial sums over the rows this rank owns, then one reduce-scatter by query slice
            s_lo, s_hi = self.offset, self.offset + self.prepared.n
            local = (i >= s_lo) & (i < s_hi)
            w = local.to(torch.float32) * (1.0 / k_mix if weights is None else weights.to(torch.float32))
            full = torch.zeros((self.world * chunk, dim), dtype=torch.float32, device=self.device)
            full[:T] = ops.gather_mix(self.synth, (i - s_lo).clamp_(0, max(self.prepared.n - 1, 0)), w)
            mine = torch.empty((chunk, dim), dtype=torch.float32, device=self.device)
            dist.reduce_scatter_tensor(mine, full, group=self.group)
            part = mine[:hi - lo]
        if gather == "slice":
            return ShardedMatch(d, i_all, part, (lo, hi), d64)
        padded = torch.zeros((chunk, dim), dtype=torch.float32, device=self.device)
        padded[:hi - lo] = part
        allf = torch.empty((self.world * chunk, dim), dtype=torch.float32, device=self.device)
        dist.all_gather_into_tensor(allf, padded, group=self.group)
        return ShardedMatch(d, i_all, allf[:T], (0, T), d64)
