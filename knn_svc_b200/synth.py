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
