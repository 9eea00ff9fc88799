"""Seeded synthetic inputs for the matcher hot path.

Nothing here is product code: it only makes reproducible feature frames, f0
tracks and harmonic pools of the shapes the reference's matcher consumes
(`ddsp_prematch_dataset.py:1163-1168`: matching_list [Np,1024], matching_f0 [Np],
harmonics [Np,49]) so that tests, golden fixtures and bench.py all draw the
same numbers.  numpy's legacy `RandomState` is used because its stream is
stable across numpy versions and machines.
"""
from __future__ import annotations

import numpy as np

D_WAVLM = 1024
N_HARM = 49
HOP = 320
SR = 16000


def randn_frames(n: int, d: int = D_WAVLM, seed: int = 0) -> np.ndarray:
    """i.i.d. N(0,1) frames, fp32 — the throughput configs (cfg 3/4)."""
    return np.random.RandomState(seed).standard_normal((n, d)).astype(np.float32)


def ar1_frames(n: int, d: int = D_WAVLM, seed: int = 0, rho: float = 0.98,
               mean_scale: float = 3.0, reset_every: int = 0,
               mean_seed: int = 12345) -> np.ndarray:
    """WavLM-like frames: AR(1) per dimension plus a large shared mean vector.

    Generator G of SURVEY.md §8d.  `reset_every` > 0 redraws the AR state every
    that many frames; consecutive frames across a reset sit at cosine distance
    ~0.1, so `2*dist >= 0.08` and the sticky branch of the greedy re-selection
    (`lib_ongaku_test.py:325-332`) fires at a known frame.
    The shared mean is drawn from `mean_seed` so that query and pool (different
    `seed`) share it, as two recordings of WavLM features do.
    """
    rs = np.random.RandomState(seed)
    mean = mean_scale * np.random.RandomState(mean_seed).standard_normal(d)
    eps = rs.standard_normal((n, d))
    x = np.empty((n, d), dtype=np.float64)
    s = np.sqrt(1.0 - rho * rho)
    state = rs.standard_normal(d)
    for t in range(n):
        if reset_every and t > 0 and t % reset_every == 0:
            state = rs.standard_normal(d)
        else:
            state = rho * state + s * eps[t]
        x[t] = state
    return (x + mean[None, :]).astype(np.float32)


def f0_track(n: int, seed: int = 0, unvoiced: float = 0.2,
             lo: float = 80.0, hi: float = 1000.0) -> np.ndarray:
    """Piecewise-smooth f0 in Hz with zeros for unvoiced frames, fp32.

    Mirrors what `get_complete_spk_pool` hands the matcher
    (`ddsp_prematch_dataset.py:373-382`; zeros below 80 Hz, `:121-128`).
    """
    rs = np.random.RandomState(seed)
    f0 = np.empty(n, dtype=np.float64)
    t = 0
    while t < n:
        seg = int(rs.randint(20, 120))
        a, b = np.exp(rs.uniform(np.log(lo * 1.2), np.log(hi * 0.8), size=2))
        b = a * np.clip(b / a, 0.8, 1.25)
        m = min(seg, n - t)
        f0[t:t + m] = np.linspace(a, b, seg)[:m] * (1.0 + 0.01 * rs.standard_normal(m))
        if rs.uniform() < unvoiced * 2.0:
            z = min(m, int(rs.randint(5, max(6, seg // 2))))
            f0[t:t + z] = 0.0
        t += m
    f0[f0 < 80.0] = 0.0
    return f0.astype(np.float32)


def harmonics_pool(n: int, h: int = N_HARM, seed: int = 0) -> np.ndarray:
    """Harmonic amplitudes [n,h] = 0.0108*|N(0,1)|/h, smoothed in time (cfg 2)."""
    rs = np.random.RandomState(seed)
    a = np.abs(rs.standard_normal((n, h)))
    for t in range(1, n):
        a[t] = 0.9 * a[t - 1] + 0.1 * a[t]
    a = 0.0108 * a / np.arange(1, h + 1)[None, :]
    return a.astype(np.float32)


def ar1_frames_device(n: int, d: int = D_WAVLM, seed: int = 0, device="cuda", seg_len: int = 500, rho: float = 0.98,
                      mean_scale: float = 3.0, mean_seed: int = 12345, out=None):
    """Generator G of SURVEY.md §8d at dataset scale, on the GPU (torch's device generator; the
    numbers differ from `ar1_frames`, the process is the same): independent AR(1) runs of
    `seg_len` frames per dimension — utterances of a pool (500 frames) or the ~200-frame stretches
    between hard resets of a query — plus the shared mean vector 3*N(0,1)^d that query and pool
    have in common (same `mean_seed`).  Fills and returns a [n, d] fp32 tensor."""
    import torch
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(mean_seed)
    mean = mean_scale * torch.randn((d,), device=dev, generator=g)
    g.manual_seed(seed)
    x = torch.empty((n, d), dtype=torch.float32, device=dev) if out is None else out
    assert x.shape == (n, d)
    s = float(np.sqrt(1.0 - rho * rho))
    n_full = n // seg_len
    # segments are advanced in lock step, a few thousand at a time (82 MB of state per 20k segments)
    for a in range(0, n_full, 8192):
        b = min(n_full, a + 8192)
        view = x[a * seg_len:b * seg_len].view(b - a, seg_len, d)
        state = torch.randn((b - a, d), device=dev, generator=g)
        for t in range(seg_len):
            if t:
                state = rho * state + s * torch.randn((b - a, d), device=dev, generator=g)
            view[:, t] = state + mean
    tail = n - n_full * seg_len
    if tail:
        state = torch.randn((d,), device=dev, generator=g)
        for t in range(tail):
            if t:
                state = rho * state + s * torch.randn((d,), device=dev, generator=g)
            x[n_full * seg_len + t] = state + mean
    return x
