"""CPU oracle for the kNN-SVC matcher hot path.  TEST INFRASTRUCTURE ONLY.

This module restates, in plain numpy, the arithmetic of the reference's matcher
path.  It exists to CHECK the CUDA path; nothing in the product
(`knn_svc_b200/`) imports it.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may call it.

Parity status: the reference ships no tests or golden vectors for this path
(SURVEY.md §4), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF,
generated in the build container by `tests/golden/make_golden.py` (which imports
the reference's own modules from /root/reference) and committed under
`tests/golden/*.npz`.  `tests/test_oracle_golden.py` checks every function here
against those fixtures.

The reference's arithmetic lives in PyTorch (`torch.cdist`, `topk`, `sort`,
`median`, `optim.Adam(amsgrad=True)`, `F.interpolate(bicubic)`, `cumsum`, `sin`);
pinned by the reference's pyproject.toml as torch ^2.9.0 (container: 2.11.0).
Each function cites the reference file:line it follows.  Everything is evaluated
in float64 unless a dtype is stated: the reference's real inference path is
float64 and its fp32 path deviates from the fp64 value by < 1e-6.
"""
from __future__ import annotations

import math

import numpy as np

# --------------------------------------------------------------------------- K1


def row_norms(x: np.ndarray) -> np.ndarray:
    """`torch.norm(x, p=2, dim=-1)` — lib_ongaku_test.py:150-151."""
    x = np.asarray(x, dtype=np.float64)
    return np.sqrt(np.einsum("ij,ij->i", x, x))


def cosine_dist(src: np.ndarray, pool: np.ndarray, direct: bool | None = None) -> np.ndarray:
    """Cosine distance matrix [T,Np] — lib_ongaku_test.py:148-175 (and the
    unchunked twin ddsp_matcher.py:213-221).

    The reference recovers the dot product from a Euclidean distance,
    dot = (-cdist^2 + |s|^2 + |p|^2)/2, then dist = 1 - dot/(|s||p|).
    `torch.cdist` evaluates |s-p|^2 as |s|^2+|p|^2-2 s.p (clamped at 0) when
    either side has more than 25 rows and as sum((s-p)^2) otherwise;
    `direct` selects the form (None = torch's rule).  The 20-row chunking of the
    reference does not change any value, only how many rows one call sees.
    """
    s = np.asarray(src, dtype=np.float64)
    p = np.asarray(pool, dtype=np.float64)
    ns, npn = row_norms(s), row_norms(p)
    if direct is None:
        direct = s.shape[0] <= 25 and p.shape[0] <= 25
    if direct:
        d2 = ((s[:, None, :] - p[None, :, :]) ** 2).sum(-1)
    else:
        d2 = np.maximum(ns[:, None] ** 2 + npn[None, :] ** 2 - 2.0 * (s @ p.T), 0.0)
    dot = (-d2 + ns[:, None] ** 2 + npn[None, :] ** 2) / 2.0
    return 1.0 - dot / (ns[:, None] * npn[None, :])


# --------------------------------------------------------------------------- K2


def topk_smallest(dists: np.ndarray, k: int):
    """`dists.topk(k, dim=-1, largest=False)` — ddsp_prematch_dataset.py:1203,
    ddsp_matcher.py:554.  Ascending; ties broken by lower index (torch's own
    tie order is unspecified, so tests mask tied rows)."""
    order = np.argsort(dists, axis=-1, kind="stable")[:, :k]
    return order.astype(np.int64), np.take_along_axis(dists, order, axis=-1)


def knn(query: np.ndarray, pool: np.ndarray, k: int = 32, mask_lo=None, mask_hi=None):
    """HOT LOOP A — ddsp_prematch_dataset.py:1196-1206: chunk-20 distance + top-32.
    With mask_lo/mask_hi ([T] each) the columns [mask_lo[t], mask_hi[t]) of row t are set to
    distance 1 before the top-k — the offline prematch's self-utterance rule,
    `dists[:, start_index:end_index] = 1`, ddsp_prematch_dataset.py:1608-1632."""
    idx, val = [], []
    for a in range(0, len(query), 20):
        d = cosine_dist(query[a:a + 20], pool, direct=False if len(pool) > 25 else None)
        if mask_lo is not None:
            for r in range(d.shape[0]):
                d[r, int(mask_lo[a + r]):int(mask_hi[a + r])] = 1.0
        i, v = topk_smallest(d, k)
        idx.append(i)
        val.append(v)
    return np.concatenate(idx, 0), np.concatenate(val, 0)


def kth_gap(query: np.ndarray, pool: np.ndarray, k: int) -> np.ndarray:
    """Distance gap between the k-th and (k+1)-th neighbour per row (the north
    star's tie criterion: indices must match wherever this exceeds 1e-5)."""
    _, v = knn(query, pool, k + 1)
    return v[:, k] - v[:, k - 1]


# --------------------------------------------------------------------------- K3


def gather_mix(pool: np.ndarray, idx: np.ndarray, weights: np.ndarray | None = None) -> np.ndarray:
    """sum_k w[t,k]*pool[idx[t,k]] — ddsp_prematch_dataset.py:1348,1358,1364,
    1435,1444,1446; ddsp_matcher.py:578 (weights None = plain mean)."""
    g = np.asarray(pool, dtype=np.float64)[idx]
    if weights is None:
        return g.mean(1)
    return (g * np.asarray(weights, dtype=np.float64)[..., None]).sum(1)


def uniform_weights(t: int, k: int) -> np.ndarray:
    """`process_weight(ones, "sum_to_1_geq")` = softmax of equal logits —
    ddsp_prematch_dataset.py:426-428, :1361."""
    return np.full((t, k), 1.0 / k)


# --------------------------------------------------------------------------- K4 / a6


def lower_median(x: np.ndarray, axis: int = -1) -> np.ndarray:
    """`torch.median`: for an even count it returns the LOWER middle element."""
    s = np.sort(x, axis=axis)
    return np.take(s, (x.shape[axis] - 1) // 2, axis=axis)


def shift_f0(query_f0: np.ndarray, pool_f0: np.ndarray) -> np.ndarray:
    """Log-domain median shift of the voiced query f0 — ddsp_prematch_dataset.py:1224-1233.
    fp32, as the reference (CPU float tensors)."""
    q = np.asarray(query_f0, dtype=np.float32)
    p = np.asarray(pool_f0, dtype=np.float32)
    mq = lower_median(np.log(q[q != 0]))
    mp = lower_median(np.log(p[p != 0]))
    out = q.copy()
    out[q != 0] = np.exp(np.log(q[q != 0]) + mp - mq).astype(np.float32)
    return out


def f0_keys(expected_f0: np.ndarray, f0_list: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """|log2(f0_cand+1e-5) - log2(f0_query+1e-5)| in fp32 — ddsp_prematch_dataset.py:974."""
    e = np.asarray(expected_f0, dtype=np.float32)
    f = np.asarray(f0_list, dtype=np.float32)[idx]
    return np.abs(np.log2(f + np.float32(1e-5)) - np.log2(e[:, None] + np.float32(1e-5))).astype(np.float32)


def sort_by_f0_compatibility(expected_f0, f0_list, idx):
    """Stable ascending re-rank of the candidates by f0 distance —
    ddsp_prematch_dataset.py:954-997 (`torch.sort(stable=True)` then gather)."""
    keys = f0_keys(expected_f0, f0_list, idx)
    order = np.argsort(keys, axis=1, kind="stable")
    return np.take_along_axis(np.asarray(idx), order, axis=1)


# --------------------------------------------------------------------------- K5


def _cosd_direct(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    return cosine_dist(x, y, direct=True)


def knn_with_concat_cost(idx, src, tgt, shifted_src_f0=None, tgt_f0=None, concat_weight=0.2,
                         return_costs: bool = False):
    """Greedy sequential re-selection — lib_ongaku_test.py:270-369.

    Row 0 is kept.  For each later frame the candidate set is the frame's own K
    neighbours plus (previous selection + 1, clamped to the last pool row);
    cost = w * lower_median_over_prev(concat cost) + matching cost [+ f0 cost];
    keep the K cheapest, ascending.  All distances are the direct-form cosine
    distance (the calls see <= 8 rows).  In the f0 branch the
    assignment `concat_weight = 0` sticks for all later frames (SURVEY D6,
    lib_ongaku_test.py:332).  With `return_costs` also returns the sorted total
    costs [T,2K] and the candidates in that order (row 0 zeros) so tests can
    mask legitimately tied rows (duplicate candidates tie harmlessly).
    """
    idx = np.asarray(idx, dtype=np.int64)
    src = np.asarray(src, dtype=np.float64)
    tgt = np.asarray(tgt, dtype=np.float64)
    t_len, k = idx.shape
    n_pool = len(tgt)
    use_f0 = shifted_src_f0 is not None
    if use_f0:
        assert tgt_f0 is not None
        lsrc = np.log2(np.asarray(shifted_src_f0, dtype=np.float64) + 1e-5)
        ltgt = np.log2(np.asarray(tgt_f0, dtype=np.float64) + 1e-5)
    w = float(concat_weight)
    out = np.empty_like(idx)
    out[0] = idx[0]
    costs = np.zeros((t_len, 2 * k))
    cands = np.zeros((t_len, 2 * k), dtype=np.int64)
    for i in range(1, t_len):
        prev = out[i - 1]
        extra = np.minimum(prev + 1, n_pool - 1)
        cand = np.concatenate([idx[i], extra])
        c_rows = tgt[cand]
        matching = _cosd_direct(src[i][None], c_rows)            # [1,2K]
        concat = _cosd_direct(tgt[prev], c_rows)                  # [K,2K]
        base = _cosd_direct(src[i - 1][None], src[i][None])[0, 0] * 2.0
        if use_f0:
            pitch = np.abs(ltgt[cand][None] - lsrc[i])
            if base < 0.08:
                concat[concat < 5.0 * base] = 0.0
            else:
                w = 0.0
            total = w * lower_median(concat, axis=0)[None] + matching + pitch
        else:
            m = concat > base
            concat[m] = 1.5 * concat[m] - base
            total = w * lower_median(concat, axis=0)[None] + matching
        order = np.argsort(total[0], kind="stable")
        out[i] = cand[order[:k]]
        costs[i] = total[0][order]
        cands[i] = cand[order]
    if return_costs:
        return out, costs, cands
    return out


# --------------------------------------------------------------------------- K6


def softmax_rows(theta: np.ndarray) -> np.ndarray:
    """`process_weight(.., "sum_to_1_geq")` = softmax over the K candidates —
    ddsp_prematch_dataset.py:426-428."""
    z = theta - theta.max(1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(1, keepdims=True)


def _neighbour_rows(idx: np.ndarray, synth: np.ndarray):
    """S_i = synth[clamp(idx+i, 0, Np-1)], i in {-1,0,1} — ddsp_prematch_dataset.py:587-595."""
    n = len(synth)
    return [synth[np.clip(idx + i, 0, n - 1)] for i in (-1, 0, 1)]


def smoothness_loss(weights: np.ndarray, rows, scale: float) -> float:
    """loss = mean_t scale*mean_d (E_-1[t+1]-E_0[t])^2 + mean_t scale*mean_d (E_0[t+1]-E_+1[t])^2
    — ddsp_prematch_dataset.py:613-636 with wavlm_phase_mae (:460, scale 0.1)
    or phase_mae (:449-457, scale 1000)."""
    e = [(r * weights[..., None]).sum(1) for r in rows]
    r1 = e[0][1:] - e[1][:-1]
    r2 = e[1][1:] - e[2][:-1]
    return float(scale * (r1 ** 2).mean(-1).mean() + scale * (r2 ** 2).mean(-1).mean())


def compute_weight(idx, synth, scale: float, max_iters: int = 100000, return_info: bool = False,
                   amp_ratio=None):
    """Adam(amsgrad) fit of per-frame softmax mixing weights —
    compute_wavlm_weight ddsp_prematch_dataset.py:574-680 (scale 0.1) and
    compute_extended_weight :807-924 (scale 1000; its `scaling_factors`
    parameter multiplies by tanh(.)*0 + 1 == 1 and is inert, :836-837).

    Control flow per iteration t (:613-670): evaluate the loss at the current
    logits; if t % 100 == 1 and |min_loss - converge_min_loss| < 1e-5 stop, else
    converge_min_loss = min_loss; if loss < min_loss snapshot the (pre-step)
    logits; stop after 1000 non-improving iterations; Adam step.  Logits and the
    Adam state are fp32 (torch parameters); the loss is evaluated in the dtype
    of `synth` promoted with fp32 weights (fp64 on the real path), and its
    gradient is cast to fp32 before the softmax backward — mirrored here.
    The AMSGrad update follows torch.optim.Adam's documented single-tensor form.

    `amp_ratio` [T,K] (compute_weight_with_amp, :684-803): every gathered row is multiplied
    by amp_ratio[t,k] for all three neighbour offsets (:713) before the same loop runs.
    """
    idx = np.asarray(idx, dtype=np.int64)
    synth = np.asarray(synth, dtype=np.float64)
    t_len, k = idx.shape
    d = synth.shape[-1]
    rows = _neighbour_rows(idx, synth)
    if amp_ratio is not None:
        amp = np.asarray(amp_ratio, dtype=np.float64)
        assert amp.shape == idx.shape                                     # :687
        rows = [r * amp[..., None] for r in rows]
    theta = np.zeros((t_len, k), dtype=np.float32)
    m = np.zeros_like(theta)
    v = np.zeros_like(theta)
    vmax = np.zeros_like(theta)
    lr, b1, b2, eps = 0.1, 0.9, 0.999, 1e-8
    min_loss = 20000.0
    converge_min_loss = 20000.0
    best = theta.copy()
    since_improve = 0
    alpha = 2.0 * scale / (max(t_len - 1, 1) * d)
    stop_t = max_iters
    for t in range(max_iters):
        w32 = softmax_rows(theta.astype(np.float32)).astype(np.float32)
        w = w32.astype(np.float64)
        e = [(r * w[..., None]).sum(1) for r in rows]
        r1 = e[0][1:] - e[1][:-1]
        r2 = e[1][1:] - e[2][:-1]
        loss = float(scale * (r1 ** 2).mean(-1).mean() + scale * (r2 ** 2).mean(-1).mean())
        if t % 100 == 1:
            if abs(min_loss - converge_min_loss) < 1e-5:
                stop_t = t
                break
            converge_min_loss = min_loss
        if loss < min_loss:
            min_loss = loss
            best = theta.copy()
            since_improve = 0
        else:
            since_improve += 1
        if since_improve >= 1000:
            stop_t = t
            break
        # dL/dE_i, then dL/dw
        g_em1 = np.zeros_like(e[0]); g_e0 = np.zeros_like(e[1]); g_ep1 = np.zeros_like(e[2])
        g_em1[1:] += alpha * r1
        g_e0[:-1] -= alpha * r1
        g_e0[1:] += alpha * r2
        g_ep1[:-1] -= alpha * r2
        g_w = (np.einsum("tkd,td->tk", rows[0], g_em1) + np.einsum("tkd,td->tk", rows[1], g_e0)
               + np.einsum("tkd,td->tk", rows[2], g_ep1)).astype(np.float32)
        g_theta = (w32 * (g_w - (g_w * w32).sum(1, keepdims=True))).astype(np.float32)
        step = t + 1
        m = (m + (g_theta - m) * np.float32(1.0 - b1)).astype(np.float32)
        v = (v * np.float32(b2) + np.float32(1.0 - b2) * g_theta * g_theta).astype(np.float32)
        vmax = np.maximum(vmax, v)
        bc1 = 1.0 - b1 ** step
        bc2_sqrt = math.sqrt(1.0 - b2 ** step)
        denom = (np.sqrt(vmax) / np.float32(bc2_sqrt) + np.float32(eps)).astype(np.float32)
        theta = (theta - np.float32(lr / bc1) * (m / denom)).astype(np.float32)
    weights = softmax_rows(best.astype(np.float32)).astype(np.float32)
    if return_info:
        return weights, {"stop_iter": stop_t, "min_loss": min_loss,
                         "loss_uniform": smoothness_loss(uniform_weights(t_len, k), rows, scale)}
    return weights


def compute_wavlm_weight(idx, synth, **kw):
    return compute_weight(idx, synth, 0.1, **kw)


def compute_extended_weight(idx, synth, **kw):
    return compute_weight(idx, synth, 1000.0, **kw)


def compute_weight_with_amp(idx, synth, amp_ratio=None, **kw):
    """ddsp_prematch_dataset.py:684-803 (phase_mae: 1000*MSE, :449-457)."""
    return compute_weight(idx, synth, 1000.0, amp_ratio=amp_ratio, **kw)


# --------------------------------------------------------------------------- K7


_CUBIC_A = -0.75


def _cubic_taps(t: np.ndarray):
    a = _CUBIC_A
    w0 = ((a * (t + 1) - 5 * a) * (t + 1) + 8 * a) * (t + 1) - 4 * a
    w1 = ((a + 2) * t - (a + 3)) * t * t + 1
    u = 1 - t
    w2 = ((a + 2) * u - (a + 3)) * u * u + 1
    w3 = ((a * (u + 1) - 5 * a) * (u + 1) + 8 * a) * (u + 1) - 4 * a
    return w0, w1, w2, w3


def upsample_bicubic(x: np.ndarray, factor: int) -> np.ndarray:
    """`F.interpolate(mode="bicubic")` along time, align_corners=False, A=-0.75,
    border-clamped taps — ddsp_prematch_dataset.py:131-143.  x [B,T,H] -> [B,T*factor,H]."""
    x = np.asarray(x, dtype=np.float64)
    t_len = x.shape[1]
    j = np.arange(t_len * factor, dtype=np.float64)
    pos = (j + 0.5) / factor - 0.5
    i0 = np.floor(pos)
    frac = pos - i0
    taps = _cubic_taps(frac)
    out = np.zeros((x.shape[0], t_len * factor, x.shape[2]))
    for o, w in zip((-1, 0, 1, 2), taps):
        ii = np.clip(i0.astype(np.int64) + o, 0, t_len - 1)
        out += x[:, ii, :] * w[None, :, None]
    return out


def wrapped_phase(f0: np.ndarray, sr: int, hop: int) -> np.ndarray:
    """fp64 running phase of the nearest-upsampled f0, wrapped to one cycle and
    cast to fp32 — ddsp_prematch_dataset.py:167,194-196 / hifigan/ddsp_models_f0.py:344-351.
    f0 [B,T] -> [B,T*hop] float32 radians in [-pi, pi]."""
    up = np.repeat(np.asarray(f0, dtype=np.float32), hop, axis=1).astype(np.float64)
    ph = np.cumsum(up / sr, axis=1)
    return (2.0 * math.pi * (ph - np.rint(ph))).astype(np.float32)


def get_bulk_dsp_choral(f0: np.ndarray, amp: np.ndarray, sample_rate: int = 16000, hop_size: int = 320) -> np.ndarray:
    """Additive harmonic bank — ddsp_prematch_dataset.py:165-208.
    f0 [B,T,1], amp [B,T,H] -> [B,T*hop,1] float32:
    sum_h sin(h*phi) * amp_up_h * ((h*f0_up < sr/2) + 1e-7)."""
    f0 = np.asarray(f0, dtype=np.float32)
    amp = np.asarray(amp, dtype=np.float32)
    n_h = amp.shape[-1]
    phi = wrapped_phase(f0[..., 0], sample_rate, hop_size)                   # fp32 [B,N]
    h = np.arange(1, n_h + 1, dtype=np.float32)
    args = (phi[..., None] * h[None, None, :]).astype(np.float32)            # fp32 product, as torch
    f0_up = np.repeat(f0[..., 0], hop_size, axis=1)
    mask = ((f0_up[..., None] * h[None, None, :]).astype(np.float32) < sample_rate / 2).astype(np.float64) + 1e-7
    amp_up = upsample_bicubic(amp, hop_size)
    sig = (np.sin(args.astype(np.float64)) * amp_up * mask).sum(-1, keepdims=True)
    return sig.astype(np.float32)


def f0_sinusoid(f0: np.ndarray, sample_rate: int = 16000, hop_size: int = 320) -> np.ndarray:
    """Single f0 sinusoid — hifigan/ddsp_models_f0.py:344-352.  f0 [B,T,1] -> [B,1,T*hop]."""
    phi = wrapped_phase(np.asarray(f0, dtype=np.float32)[..., 0], sample_rate, hop_size)
    return np.sin(phi.astype(np.float64)).astype(np.float32)[:, None, :]


# --------------------------------------------------------------------------- a11


def parse_post_opt(post_opt: str) -> float:
    """"post_opt_0.2" -> 0.2, "..._extra" -> 0.3, anything else -> -1 (off) —
    ddsp_prematch_dataset.py:1273-1279."""
    tail = post_opt.split("_")[-1]
    try:
        return float(tail)
    except ValueError:
        return 0.3 if tail == "extra" else -1.0


def match_utterance(query, query_f0, pool, pool_f0, harmonics, post_opt="no_post_opt",
                    ckpt_type="mix", synth=None):
    """Tensor-level body of match_at_inference_time for one query utterance —
    ddsp_prematch_dataset.py:1180-1451 (the file/WavLM/pyworld front end,
    :1109-1168, is outside the hot path).  Returns a dict with the matched
    features [T,D] fp32, shifted f0 [T], mixed harmonics [T,H] (mix only) and the
    intermediate index sets."""
    synth = pool if synth is None else synth
    nbrs, _ = knn(query, pool, 32)                                   # :1196-1206
    shifted = shift_f0(query_f0, pool_f0)                            # :1224-1233
    cw = parse_post_opt(post_opt)
    idx = nbrs[:, :4].copy()                                         # :1246 (topk ignored, D4)
    if cw != -1:
        idx = knn_with_concat_cost(idx, query, pool, concat_weight=cw)   # :1295
    if "no_post_opt" not in post_opt:
        w = compute_wavlm_weight(idx, synth)                         # :1357
    else:
        w = uniform_weights(*idx.shape)                              # :1361
    feats = gather_mix(synth, idx, w).astype(np.float32)             # :1358,:1364
    out = {"nearest_nbrs": nbrs, "wavlm_indices": idx, "wavlm_weights": np.asarray(w),
           "out_feats": feats, "shifted_f0": shifted}
    prio = sort_by_f0_compatibility(shifted, pool_f0, nbrs)          # :1377
    idx_h = prio[:, :4].copy()                                       # :1398
    if cw != -1:
        idx_h = knn_with_concat_cost(idx_h, query, pool, shifted, pool_f0, concat_weight=cw)  # :1414
    out["f0_priority"] = prio
    out["harm_indices"] = idx_h
    if "wavlm_only" not in ckpt_type and "no_harm_no_amp" not in ckpt_type:
        if "no_post_opt" not in post_opt:
            wh = compute_extended_weight(idx_h, harmonics)           # :1441
            out["harm_weights"] = wh
            out["harmonics"] = gather_mix(harmonics, idx_h, wh)      # :1444
        else:
            out["harmonics"] = gather_mix(harmonics, idx_h, None)    # :1446
    return out


# --------------------------------------------------------------------------- §8f rank 3: pool-builder tensor ops


def layer_mix(feats: np.ndarray, weights: np.ndarray) -> np.ndarray:
    """`(feats * weights[:, None]).sum(dim=0)` — ddsp_prematch_dataset.py:349-350.
    feats [L,T,D], weights [L,1] or [L] -> [T,D] (float64, as the reference's float64
    weighting produces)."""
    f = np.asarray(feats, dtype=np.float64)
    w = np.asarray(weights, dtype=np.float64).reshape(-1)
    return np.einsum("ltd,l->td", f, w)


def stft_magnitude(x: np.ndarray, n_frames: int | None = None, n_fft: int = 400, hop: int = 320) -> np.ndarray:
    """`torchaudio.transforms.Spectrogram(n_fft=400, hop_length=320, center=True, power=1)`
    (periodic Hann window, reflect padding, one-sided) followed by `.T[:, :-1]` and the crop to
    the WavLM frame count — ddsp_prematch_dataset.py:326, :361-363.  x [N] -> [frames, n_fft/2]."""
    x = np.asarray(x, dtype=np.float64)
    pad = n_fft // 2
    xp = np.pad(x, (pad, pad), mode="reflect")
    frames = 1 + len(x) // hop
    win = 0.5 - 0.5 * np.cos(2.0 * math.pi * np.arange(n_fft) / n_fft)
    seg = np.stack([xp[i * hop:i * hop + n_fft] for i in range(frames)]) * win[None, :]
    mag = np.abs(np.fft.rfft(seg, axis=1))[:, :-1]
    return mag if n_frames is None else mag[:n_frames]


def interp_linear(spec: np.ndarray, factor: int = 8) -> np.ndarray:
    """`F.interpolate(spec[None], scale_factor=8, mode='linear')` along the frequency axis
    (align_corners=False) — ddsp_prematch_dataset.py:395.  [T,S] -> [T,S*factor]; torch
    evaluates the source position and both weights in the tensor dtype (fp32)."""
    spec = np.asarray(spec, dtype=np.float32)
    s_in = spec.shape[1]
    j = np.arange(s_in * factor, dtype=np.float32)
    src = np.maximum(np.float32(1.0 / factor) * (j + np.float32(0.5)) - np.float32(0.5), np.float32(0)).astype(np.float32)
    i0 = np.floor(src).astype(np.int64)
    i1 = np.minimum(i0 + 1, s_in - 1)
    lam1 = (src - i0.astype(np.float32)).astype(np.float32)
    lam0 = (np.float32(1) - lam1).astype(np.float32)
    # torch's CPU kernel evaluates lam0*x0 + lam1*x1 as fma(lam0, x0, round(lam1*x1)) (measured:
    # bit-exact on the fixture); the device kernel uses the same contraction.
    hi = (lam1[None, :] * spec[:, i1]).astype(np.float32).astype(np.float64)
    return (lam0[None, :].astype(np.float64) * spec[:, i0].astype(np.float64) + hi).astype(np.float32)


def harmonic_amplitudes(spec: np.ndarray, f0: np.ndarray, n_harm: int = 49, sr: int = 16000) -> np.ndarray:
    """Amplitudes of the first 49 harmonics read off the x8-interpolated magnitude spectrum —
    ddsp_prematch_dataset.py:391-404.  spec [T,S] (S=200), f0 [T] -> [T,49] fp32:
    bin = round(clamp(f0*h*2*(8S)/sr, max=8S)) into the spectrum padded with one 0;
    unvoiced frames (f0 == 0): harmonic 1 = max of the raw spectrum row, others 0; all x0.0108."""
    spec = np.asarray(spec, dtype=np.float32)
    f0 = np.asarray(f0, dtype=np.float32)
    up = interp_linear(spec, 8)
    n_up = up.shape[1]
    h = np.arange(1, n_harm + 1, dtype=np.float32)
    mh = (f0[:, None] * h[None, :]).astype(np.float32)                       # :391
    pos = (((mh * np.float32(2)).astype(np.float32) * np.float32(n_up)).astype(np.float32) / np.float32(sr)).astype(np.float32)
    bins = np.rint(np.minimum(pos, np.float32(n_up))).astype(np.int64)       # :397 (round half to even)
    padded = np.concatenate([up, np.zeros((len(up), 1), np.float32)], axis=1)   # F.pad(.., (0, 1))
    out = np.take_along_axis(padded, bins, axis=1)
    unvoiced = f0 == 0
    out[unvoiced, 1:] = 0                                                    # :401
    out[unvoiced, 0] = spec.max(1)[unvoiced]                                 # :402
    return (np.float32(0.0108) * out).astype(np.float32)                     # :404


def amp_ratio(spec_query: np.ndarray, spec_pool: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """L1-norm ratio of the utterance's own spectrum rows to the gathered candidates' —
    ddsp_prematch_dataset.py:1672-1675.  [T,S], [Np,S], [T,K] -> [T,K]."""
    orig = np.abs(np.asarray(spec_query, dtype=np.float64)).sum(1)
    knn_l1 = np.abs(np.asarray(spec_pool, dtype=np.float64)).sum(1)[idx]
    return orig[:, None] / (knn_l1 + 1e-5)


# --------------------------------------------------------------------------- §8f rank 1: offline prematch


def half_round(x: np.ndarray) -> np.ndarray:
    """`.half().float()` — ddsp_prematch_dataset.py:1510, :1567, :1613."""
    return np.asarray(x, dtype=np.float32).astype(np.float16).astype(np.float32)


def prematch_utterance(start: int, end: int, matching_list: np.ndarray, f0_list: np.ndarray,
                       spec_list: np.ndarray, harmonics_list: np.ndarray, fit: bool = True, nbrs=None):
    """Body of per_spk_extract for ONE utterance occupying pool rows [start, end) —
    ddsp_prematch_dataset.py:1608-1769: masked top-32 against the speaker's own pool, f0
    re-rank, amp_ratio, compute_weight_with_amp.  `matching_list` is the `.half().float()`
    rounded pool (:1567); the query rows are the same rows (:1613)."""
    t_len = end - start
    q = matching_list[start:end]
    lo = np.full(t_len, start, np.int64)
    hi = np.full(t_len, end, np.int64)
    if nbrs is None:   # (tests pass the reference's own top-32 to compare the later stages past tied slots)
        nbrs, _ = knn(q, matching_list, 32, lo, hi)                          # :1608-1632
    prio = sort_by_f0_compatibility(f0_list[start:end], f0_list, nbrs)       # :1646
    idx = prio[:, :4].copy()                                                 # :1655
    ratio = amp_ratio(spec_list[start:end], spec_list, idx)                  # :1672-1675
    out = {"slice": (start, end), "nearest_nbrs": nbrs, "nearest_nbrs_f0_priority": prio,
           "amp_ratio": ratio.astype(np.float32)}
    if fit:
        out["harmonics_best_weight_para"] = compute_weight_with_amp(idx, harmonics_list, amp_ratio=ratio)   # :1681
    return out
