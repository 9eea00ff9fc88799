"""CPU baseline for bench.py: the reference's matcher loop re-stated with the
SAME library calls the reference makes (torch on the host cores), timed on a
bounded sample.  TEST/MEASUREMENT INFRASTRUCTURE ONLY — never on a product path.

Follows ddsp_prematch_dataset.py:1196-1206 (chunk-20 fast_cosine_dist + topk(32),
lib_ongaku_test.py:148-175), then `[:, :4]` and the gather-mean (:1246, :1348, :1364).
kind = "port": /root/reference is not present on the GPU box, so the reference's
own file cannot be imported there.
"""
from __future__ import annotations

import time

import torch


def _fast_cosine_dist(src, pool, increment=20):
    """lib_ongaku_test.py:148-175, call for call: BOTH norm vectors are computed inside the function,
    i.e. the pool norms are recomputed on every call — and the matcher calls it once per 20-row chunk
    of the query (ddsp_prematch_dataset.py:1200), so the pool is re-read for its norms every chunk."""
    src_norms_all = torch.norm(src, p=2, dim=-1)
    pool_norms = torch.norm(pool, p=2, dim=-1)
    out = []
    for a in range(0, len(src), increment):
        sn, sf = src_norms_all[a:a + increment], src[a:a + increment]
        dot = -torch.cdist(sf[None], pool[None], p=2)[0] ** 2 + sn[:, None] ** 2 + pool_norms[None] ** 2
        dot /= 2
        d = 1 - (dot / (sn[:, None] * pool_norms[None]))
        if torch.sum(torch.isnan(d)) > 0:
            raise SystemExit("containing nan")
        out.append(d)
    return torch.cat(out, dim=0)


def match_sample(query: torch.Tensor, pool: torch.Tensor, k_search: int = 32, k_mix: int = 4, increment: int = 20):
    """One pass of the reference's matcher over `query` (CPU tensors): the chunk-20 loop of
    ddsp_prematch_dataset.py:1196-1206 calling fast_cosine_dist on each chunk."""
    nbrs = []
    for a in range(0, len(query), increment):
        dists = _fast_cosine_dist(query[a:a + increment], pool)
        nbrs.append(dists.topk(k=min(k_search, pool.shape[0]), dim=-1, largest=False).indices)
    nbrs = torch.cat(nbrs, dim=0)
    idx = nbrs[:, :k_mix]
    out = pool[idx.reshape(-1)].reshape(idx.shape[0], idx.shape[1], pool.shape[1]).mean(1)
    return nbrs, out


def use_all_host_threads() -> int:
    """torchrun exports OMP_NUM_THREADS=1; the baseline is meant to use every host core it may."""
    import os
    if torch.get_num_threads() > 1:
        return torch.get_num_threads()          # torch's own default (affinity / cgroup aware) stands
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    torch.set_num_threads(max(1, min(n, 32)))   # capped: 20-row GEMM chunks do not scale past a few dozen threads
    return torch.get_num_threads()


def time_sample(n_query: int, n_pool: int, dim: int, steps: int = 1, warmup: int = 0, seed: int = 0):
    """Returns (seconds per pass, threads used)."""
    use_all_host_threads()
    g = torch.Generator().manual_seed(seed)
    q = torch.randn((n_query, dim), generator=g)
    p = torch.randn((n_pool, dim), generator=g)
    for _ in range(warmup):
        match_sample(q, p)
    t0 = time.perf_counter()
    for _ in range(steps):
        match_sample(q, p)
    return (time.perf_counter() - t0) / max(steps, 1), torch.get_num_threads()
