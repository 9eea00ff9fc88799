"""Pool-sharded kNN across the GPUs of one box.

One process per GPU.  Rank r owns pool rows [r*Np/R, (r+1)*Np/R); queries are
replicated.  Each rank runs the fused distance/top-k on its shard, the per-rank
(k distances + k global indices) per query frame are exchanged with ONE
all-gather (NCCL over NVLink on GPUs; gloo in the CPU tests of the host logic),
and every rank merges them with ties broken by the lower global index, so the
result does not depend on the shard count.  The reference has no counterpart:
its matcher is single-device (ddsp_prematch_dataset.py:1196-1206).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_pool: int, world_size: int, rank: int):
    """Contiguous, balanced split by pool frame."""
    lo = (n_pool * rank) // world_size
    hi = (n_pool * (rank + 1)) // world_size
    return lo, hi


def all_gather_topk(local_dist: torch.Tensor, local_idx: torch.Tensor, group=None):
    """[T,k] per rank -> ([R,T,k], [R,T,k]) on every rank, one collective each for
    distances and indices (packed into a single buffer when dtypes allow)."""
    world = dist.get_world_size(group)
    T, k = local_dist.shape
    # pack fp32 distances (bit pattern) and int64 indices into one int64 payload: one all-gather
    payload = torch.empty((T, 2 * k), dtype=torch.int64, device=local_dist.device)
    payload[:, :k] = local_dist.contiguous().view(torch.int32).to(torch.int64)
    payload[:, k:] = local_idx
    out = torch.empty((world * T, 2 * k), dtype=torch.int64, device=local_dist.device)
    dist.all_gather_into_tensor(out, payload, group=group)
    out = out.view(world, T, 2 * k)
    gd = out[:, :, :k].to(torch.int32).view(torch.float32)
    gi = out[:, :, k:].contiguous()
    return gd.contiguous(), gi


def merge_topk_host(gd: torch.Tensor, gi: torch.Tensor):
    """Merge rule in plain torch, (dist, idx) lexicographic — used by the gloo CPU
    test of the exchange logic; GPUs use ops.merge_topk (the CUDA kernel)."""
    R, T, k = gd.shape
    d = gd.permute(1, 0, 2).reshape(T, R * k)
    i = gi.permute(1, 0, 2).reshape(T, R * k)
    valid = i >= 0
    d = torch.where(valid, d, torch.full_like(d, float("inf")))
    order = torch.argsort(i, dim=1, stable=True)          # secondary key first
    d, i = torch.gather(d, 1, order), torch.gather(i, 1, order)
    order = torch.argsort(d, dim=1, stable=True)          # then primary key, stably
    return torch.gather(d, 1, order)[:, :k].contiguous(), torch.gather(i, 1, order)[:, :k].contiguous()


class ShardedPool:
    """This rank's slice of the target pool, prepared for the tensor-core filter."""

    def __init__(self, shard_rows: torch.Tensor, global_offset: int, group=None):
        from . import ops
        self.prepared = ops.prepare_rows(shard_rows)
        self.offset = int(global_offset)
        self.group = group

    def knn(self, query_prepared, k: int):
        from . import ops
        d, i = ops.knn_search(query_prepared, self.prepared, k, index_offset=self.offset)
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return d, i
        gd, gi = all_gather_topk(d, i, self.group)
        return ops.merge_topk(gd, gi)
