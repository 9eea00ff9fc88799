"""Drop-in replacements for the matcher functions of the reference's
`ddsp_prematch_dataset.py`, backed by the CUDA library.

    from knn_svc_b200.ddsp_prematch_dataset import (
        match_at_inference_time, get_bulk_dsp_choral, sort_by_f0_compatibility,
        compute_wavlm_weight, compute_extended_weight, process_weight)

The WavLM / pyworld front end (`get_complete_spk_pool`, reference :301-423) is a
feature PRODUCER and stays in the reference: `match_at_inference_time` calls the
module attribute `get_complete_spk_pool`, which resolves to the reference's
function when the reference is importable, or to whatever the caller assigns.
"""
from __future__ import annotations

import copy
import os

import torch
import torch.nn.functional as F

from . import ops
from .lib_ongaku_test import knn_with_concat_cost  # noqa: F401  (re-export, reference imports it here too)

DOWNSAMPLE_FACTOR = 320


def get_complete_spk_pool(*args, **kwargs):
    """Resolved lazily to the reference's pool builder (ddsp_prematch_dataset.py:301)."""
    try:
        import importlib
        ref = importlib.import_module("ddsp_prematch_dataset")
    except Exception as exc:  # pragma: no cover - depends on the deployment
        raise RuntimeError(
            "get_complete_spk_pool is the reference's WavLM/pyworld front end; put the reference on "
            "PYTHONPATH or assign knn_svc_b200.ddsp_prematch_dataset.get_complete_spk_pool") from exc
    return ref.get_complete_spk_pool(*args, **kwargs)


def process_weight(weight_para, process_type):
    """reference :426-447 (only the variant on the live path is device-agnostic torch)."""
    if process_type == "sum_to_1_geq":
        return F.softmax(weight_para, dim=1)
    if process_type == "sum_to_1":
        temp = weight_para + 1 / weight_para.shape[1]
        return temp / (torch.sum(temp, dim=1, keepdim=True) + 1e-5)
    if process_type == "geq":
        e = torch.exp(weight_para)
        return 3 * e / (e + 2) / weight_para.shape[1]
    if process_type == "none":
        return weight_para + 1 / weight_para.shape[1]
    raise NotImplementedError


def sort_by_f0_compatibility(expected_f0, f0_list, target_feature_indices):
    """Stable re-rank of each row's candidates by |log2 f0| distance — reference :954-997."""
    if len(expected_f0) != len(target_feature_indices):
        raise AssertionError("expected_f0 and indices must have the same number of frames")
    return ops.f0_rerank(expected_f0, f0_list, target_feature_indices)


def compute_wavlm_weight(target_feature_indices, synth_set, process_type="sum_to_1_geq", utt_offsets=None):
    """Per-frame mixing weights minimising the neighbour-frame smoothness loss —
    reference :574-680 (Adam amsgrad on softmax logits, loss 0.1*MSE).  `utt_offsets`
    (extension) fits a concatenated batch of utterances independently in one launch."""
    if process_type != "sum_to_1_geq":
        raise NotImplementedError("only sum_to_1_geq is on the reference's live path")
    return ops.weight_fit(target_feature_indices, synth_set, 0.1, utt_offsets=utt_offsets)


def compute_extended_weight(target_feature_indices, synth_set, process_type="sum_to_1_geq", factors=[1],
                            utt_offsets=None):
    """Same fit on the harmonic amplitudes, loss 1000*MSE — reference :807-924.
    `factors` must be [1] (the only value the reference passes, :1440)."""
    if process_type != "sum_to_1_geq" or list(factors) != [1]:
        raise NotImplementedError("only sum_to_1_geq with factors=[1] is on the reference's live path")
    return ops.weight_fit(target_feature_indices, synth_set, 1000.0, utt_offsets=utt_offsets)


def compute_weight_with_amp(target_feature_indices, synth_set, process_type="sum_to_1_geq", amp_ratio=None,
                            utt_offsets=None):
    """Training-time variant of the fit (offline prematch): every candidate row is scaled by
    amp_ratio[t,k] before mixing, loss 1000*MSE — reference :684-803, called at :1681."""
    if process_type != "sum_to_1_geq":
        raise NotImplementedError("only sum_to_1_geq is on the reference's live path")
    if amp_ratio is not None:
        assert amp_ratio.shape == target_feature_indices.shape                      # :687
    return ops.weight_fit(target_feature_indices, synth_set, 1000.0, utt_offsets=utt_offsets, amp_ratio=amp_ratio)


def get_bulk_dsp_choral(f0, amp, sample_rate=16000, hop_size=320):
    """Additive harmonic bank — reference :165-208.  f0 [B,T,1], amp [B,T,H] -> [B,T*hop,1]."""
    assert f0.device == amp.device, [f0.device, amp.device]
    return ops.harmonic_bank(f0[..., 0], amp, sample_rate, hop_size)[..., None]


def f0_sinusoid(f0, sample_rate=16000, hop_size=320):
    """Single f0 sinusoid of hifigan/ddsp_models_f0.py:344-352.  f0 [B,T,1] -> [B,1,T*hop]."""
    return ops.harmonic_bank(f0[..., 0], None, sample_rate, hop_size)[:, None, :]


def shift_query_f0(query_f0, matching_f0):
    """Log-domain median shift of the voiced frames — reference :1224-1233 (torch, as there)."""
    query_f0_median = torch.median(torch.log(query_f0[query_f0 != 0]))
    matching_f0_median = torch.median(torch.log(matching_f0[matching_f0 != 0]))
    shifted = copy.deepcopy(query_f0)
    shifted[query_f0 != 0] = torch.exp(torch.log(query_f0[query_f0 != 0]) + matching_f0_median - query_f0_median)
    return shifted


def _lower_median_rows(x: torch.Tensor, valid: torch.Tensor) -> torch.Tensor:
    """torch.median (lower median) of the valid entries of each row of a padded [U, L] batch."""
    big = torch.finfo(x.dtype).max
    srt = torch.sort(torch.where(valid, x, torch.full_like(x, big)), dim=1).values
    cnt = valid.sum(1)
    pos = ((cnt - 1).clamp_min(0) // 2)[:, None]
    return srt.gather(1, pos)[:, 0]


def shift_query_f0_batched(f0_list, matching_f0_median: torch.Tensor):
    """`shift_query_f0` (reference :1224-1233) for a batch of utterances with a handful of batched
    torch ops instead of ~10 small ones per utterance.  Returns one concatenated [sum T] tensor on
    the device of `matching_f0_median`."""
    dev = matching_f0_median.device
    lens = [int(len(f)) for f in f0_list]
    U, L = len(lens), max(lens + [1])
    if U and all(f.device.type == "cpu" for f in f0_list):
        # f0 tracks live on the host (reference :373-382): pad there, ONE host-to-device copy
        pad = torch.nn.utils.rnn.pad_sequence(list(f0_list), batch_first=True).to(dev, non_blocking=True)
        if pad.shape[1] < L:
            pad = torch.nn.functional.pad(pad, (0, L - pad.shape[1]))
    else:
        pad = torch.zeros((U, L), dtype=f0_list[0].dtype if U else torch.float32, device=dev)
        for u, f in enumerate(f0_list):
            pad[u, :lens[u]] = f.to(dev)
    voiced = pad != 0
    logf = torch.log(torch.where(voiced, pad, torch.ones_like(pad)))
    med = _lower_median_rows(logf, voiced)                      # NaN-free even for all-unvoiced rows
    shifted = torch.where(voiced, torch.exp(logf + matching_f0_median.to(pad.dtype) - med[:, None]), pad)
    # the valid entries, row by row — gathered through indices built on the host (a boolean mask would make torch
    # read the count back: a blocking device-to-host copy in the middle of the batch)
    import numpy as np
    take = np.concatenate([np.arange(n, dtype=np.int64) + u * L for u, n in enumerate(lens)]) if U else np.zeros(0, np.int64)
    return shifted.reshape(-1).index_select(0, torch.from_numpy(take).to(dev))


def parse_post_opt(post_opt: str) -> float:
    """reference :1273-1279"""
    try:
        return float(post_opt.split("_")[-1])
    except ValueError:
        return 0.3 if post_opt.split("_")[-1] == "extra" else -1


_side_streams: dict = {}


def _side_stream(device) -> "torch.cuda.Stream":
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=key)
    return _side_streams[key]


class MatchingPool:
    """Target-speaker pool resident in HBM: fp32 rows, fp16 tensor-core operand,
    norms, f0 and harmonic amplitudes (reference :1163-1168 builds the same
    concatenation on every call)."""

    def __init__(self, matching_list, synth_list, matching_f0, harmonics_synth_list, device):
        self.device = torch.device(device)
        self.matching = ops.prepare_rows(matching_list.to(self.device))
        same = synth_list is matching_list or (synth_list.shape == matching_list.shape
                                               and synth_list.data_ptr() == matching_list.data_ptr())
        self.synth = self.matching.rows if same else synth_list.to(self.device, torch.float32).contiguous()
        self.f0 = matching_f0.to(torch.float32)
        self.f0_dev = self.f0.to(self.device)
        voiced = self.f0_dev[self.f0_dev != 0]
        # lower median of the voiced log-f0 (reference :1227), computed once per pool
        self.log_f0_median = torch.median(torch.log(voiced)) if len(voiced) else torch.zeros((), device=self.device)
        self.harmonics = None if harmonics_synth_list is None else \
            harmonics_synth_list.to(self.device, torch.float32).contiguous()


def match_utterances(query_seqs, query_f0s, pool: MatchingPool, post_opt="no_post_opt", ckpt_type="mix",
                     prioritize_f0=True):
    """Tensor-level body of match_at_inference_time (reference :1180-1451) for a BATCH of query
    utterances against one pool (BASELINE cfg 5): the utterances are concatenated and every stage
    runs once over the batch — one fused kNN search, one greedy re-selection launch (one CTA per
    utterance), one weight-fit launch (one CTA per utterance), one gather-mix.  Each utterance's
    results do not depend on what else is in the batch.  Returns a list of result dicts (views
    into the batch tensors)."""
    assert prioritize_f0                                                     # reference :1375
    if not len(query_seqs):
        return []
    dev = pool.device
    # K5 and K6 run one CTA per utterance and the GPU hands CTAs out in index order: with more
    # utterances than SMs, longest-first keeps the tail of the launch short.  Results go back in
    # the caller's order (and do not depend on the order: every utterance is independent).
    order = list(range(len(query_seqs)))
    if len(order) > 1 and os.environ.get("KNNSVC_BATCH_ORDER", "longest_first") == "longest_first":
        order.sort(key=lambda u: -int(query_seqs[u].shape[0]))
    place = {u: pos for pos, u in enumerate(order)}
    caller_f0s = query_f0s
    query_seqs = [query_seqs[u] for u in order]
    query_f0s = [query_f0s[u] for u in order]
    lens = [int(q.shape[0]) for q in query_seqs]
    offs = [0]
    for n in lens:
        offs.append(offs[-1] + n)
    query = ops.prepare_rows(torch.concat([q.to(dev) for q in query_seqs], dim=0))
    _, nearest_nbrs = ops.knn_search(query, pool.matching, 32)               # :1196-1206
    shifted_f0 = shift_query_f0_batched(query_f0s, pool.log_f0_median)       # :1224-1233
    concat_weight = parse_post_opt(post_opt)
    fit = "no_post_opt" not in post_opt
    # Two independent chains follow the search — features (plain top-4 -> K5 -> K6 -> mix) and
    # harmonics (f0 re-rank -> K5 with f0 -> K6 -> mix).  For a single utterance each stage is one
    # CTA walking a serial recurrence, so the chains run on two streams and overlap.
    main = torch.cuda.current_stream(dev)
    side = _side_stream(dev) if (concat_weight != -1 or fit) else main     # nothing long to overlap without post_opt
    side.wait_stream(main)
    harm = None
    with torch.cuda.stream(side):
        prio = sort_by_f0_compatibility(shifted_f0, pool.f0_dev, nearest_nbrs)   # :1377
        idx_h = prio[:, :4].contiguous()                                         # :1398
        if concat_weight != -1:
            idx_h = ops.concat_cost_reselect(idx_h, query.rows, pool.matching.rows, shifted_f0, pool.f0_dev,
                                             concat_weight=concat_weight, utt_offsets=offs)      # :1414
        if "wavlm_only" not in ckpt_type and "no_harm_no_amp" not in ckpt_type:
            hw = compute_extended_weight(idx_h, pool.harmonics, "sum_to_1_geq", [1], utt_offsets=offs) if fit else None
            harm = ops.gather_mix(pool.harmonics, idx_h, hw)                     # :1444 / :1446
    idx_w = nearest_nbrs[:, :4].contiguous()                                 # :1246
    if concat_weight != -1:
        idx_w = ops.concat_cost_reselect(idx_w, query.rows, pool.matching.rows, concat_weight=concat_weight,
                                         utt_offsets=offs)                   # :1295
    w = compute_wavlm_weight(idx_w, pool.synth, "sum_to_1_geq", utt_offsets=offs) if fit else None   # :1357 / :1361
    out_feats = ops.gather_mix(pool.synth, idx_w, w)                         # :1358 / :1364
    main.wait_stream(side)
    for t in (prio, idx_h, harm):
        if t is not None:
            t.record_stream(main)
    # the reference hands the shifted f0 back on the device its f0 came from (the host): one copy
    f0_devs = {f.device for f in query_f0s}
    if len(f0_devs) == 1 and next(iter(f0_devs)).type == "cpu":
        shifted_host = ops.to_host_small(shifted_f0)      # (not through the copy engine: see ops.to_host_small)
    else:
        shifted_host = shifted_f0.to(next(iter(f0_devs))) if len(f0_devs) == 1 else None
    results = []
    for u in range(len(lens)):                      # u: the caller's index; its data sits at position place[u]
        a, b = offs[place[u]], offs[place[u] + 1]
        r = {"out_feats": out_feats[a:b],
             "shifted_f0": shifted_host[a:b] if shifted_host is not None else shifted_f0[a:b].to(caller_f0s[u].device),
             "wavlm_indices": idx_w[a:b], "nearest_nbrs": nearest_nbrs[a:b], "harm_indices": idx_h[a:b]}
        if harm is not None:
            r["harmonics"] = harm[a:b]
        results.append(r)
    return results


def match_utterance(query_seq, query_f0, pool: MatchingPool, post_opt="no_post_opt", ckpt_type="mix",
                    prioritize_f0=True):
    """Tensor-level body of match_at_inference_time for one query utterance (reference :1180-1451):
    a batch of one."""
    return match_utterances([query_seq], [query_f0], pool, post_opt=post_opt, ckpt_type=ckpt_type,
                            prioritize_f0=prioritize_f0)[0]


class PoolCache:
    """In-process cache of speaker pools for dataset -> dataset conversion.
    The reference rebuilds both pools — WavLM over every file of both speakers — for every
    (source speaker, target speaker) pair and force-disables its pickle cache
    (ddsp_prematch_dataset.py:1086-1087, ddsp_matcher.py:1073-1112).  Here the query-side
    feature dicts and the target side's HBM-resident `MatchingPool` (fp32 rows, fp16 operand,
    norms, f0 median) are kept per (path, duration_limit), least recently used evicted first."""

    def __init__(self, max_entries: int = 16):
        from collections import OrderedDict
        self.max_entries = max_entries
        self._d = OrderedDict()
        self.hits = 0
        self.misses = 0

    def get(self, key, build):
        if key in self._d:
            self._d.move_to_end(key)
            self.hits += 1
            return self._d[key]
        self.misses += 1
        val = build()
        self._d[key] = val
        while len(self._d) > self.max_entries:
            self._d.popitem(last=False)
        return val


def match_at_inference_time(src_wav_file, ref_wav_file, wavlm, match_weights, synth_weights, topk: int = 4,
                            device="cuda", prioritize_f0=False, ckpt_type="wavlm_only", src_dataset_path=None,
                            tgt_dataset_path=None, cache_dir=None, required_subset=None, post_opt="no_post_opt",
                            duration_limit=None, pool_cache: "PoolCache | None" = None):
    """Same call, same returns as the reference (ddsp_prematch_dataset.py:1074-1459):
    dicts keyed by query file of matched features [T,D] fp32, (mix only) mixed
    harmonics [T,49], a dict of None, and the shifted f0 [T].  `pool_cache` (extension,
    default off = the reference's behaviour) reuses pools across calls, see PoolCache."""
    if src_dataset_path is None:
        assert os.path.isfile(src_wav_file)

    def build_query():
        return get_complete_spk_pool(src_wav_file, wavlm, match_weights, synth_weights, device=device)

    def build_target():
        matching_pool, synth_pool, audio_synth_pool, spec_synth_pool, f0_pool, harmonics_synth_pool = \
            get_complete_spk_pool(ref_wav_file, wavlm, match_weights, synth_weights, device=device,
                                  duration_limit=duration_limit)
        keys = list(matching_pool)
        return MatchingPool(torch.concat([matching_pool[k] for k in keys], dim=0),
                            torch.concat([synth_pool[k] for k in keys], dim=0),
                            torch.concat([f0_pool[k] for k in keys], dim=0),
                            torch.concat([harmonics_synth_pool[k] for k in keys], dim=0), device)

    if tgt_dataset_path is None:
        assert os.path.isfile(ref_wav_file)
    if pool_cache is None:
        query_side, pool = build_query(), build_target()
    else:
        query_side = pool_cache.get(("query", str(src_wav_file), str(device)), build_query)
        pool = pool_cache.get(("target", str(ref_wav_file), duration_limit, str(device)), build_target)
    query_pool, _, _, query_spec_pool, query_f0_pool, _ = query_side
    out_feats_weighted_collection = dict()
    harmonics_out_feats_weighted_collection = dict()
    audio_out_feats_weighted_collection = dict()
    shifted_query_f0_collection = dict()
    items = [item for item in query_pool
             if required_subset is None or
             os.path.basename(item).split(".")[0] + "/" + os.path.basename(ref_wav_file) in required_subset]   # :1181
    # all query utterances of this (source, target) pair go through the matcher as ONE batch
    results = match_utterances([query_pool[i] for i in items], [query_f0_pool[i] for i in items], pool,
                               post_opt=post_opt, ckpt_type=ckpt_type, prioritize_f0=prioritize_f0)
    for item, res in zip(items, results):
        out_feats_weighted_collection[item] = res["out_feats"]
        audio_out_feats_weighted_collection[item] = None
        shifted_query_f0_collection[item] = res["shifted_f0"]
        if "harmonics" in res:
            harmonics_out_feats_weighted_collection[item] = res["harmonics"]
    if "wavlm_only" in ckpt_type or "no_harm_no_amp" in ckpt_type:
        return out_feats_weighted_collection, audio_out_feats_weighted_collection, shifted_query_f0_collection
    elif "mix" in ckpt_type:
        return (out_feats_weighted_collection, harmonics_out_feats_weighted_collection,
                audio_out_feats_weighted_collection, shifted_query_f0_collection)
    raise NotImplementedError


# ----------------------------------------------------------------------------- SURVEY §8(f)
# rank 3: the tensor arithmetic of the pool builder after WavLM; rank 1: the offline prematch.


def spk_pool_from_features(feats, x, f0, match_weights, synth_weights, device="cuda"):
    """Tensor part of get_complete_spk_pool for ONE utterance (reference :347-404): given the
    WavLM layer stack `feats` [L,T,D], the mono 16 kHz waveform `x` [N] and the f0 track
    [T or T+1], returns (matching [T,D], synth [T,D], audio frames [T,320], spec [T,200],
    f0 [T], harmonic amplitudes [T,49]) — all on `device`, fp32.  WavLM, file IO and pyworld
    (the producers of feats / x / f0) stay with the caller."""
    dev = torch.device(device)
    feats = feats.to(dev)
    mw = match_weights.detach().cpu().double().reshape(-1).numpy()
    sw = synth_weights.detach().cpu().double().reshape(-1).numpy()
    matching, synth = ops.layer_mix(feats, mw, sw)                                   # :349-350
    T = matching.shape[0]
    x = x.to(dev).reshape(-1).float()
    assert len(x) >= DOWNSAMPLE_FACTOR * T                                           # :354
    audio = x[:DOWNSAMPLE_FACTOR * T].reshape(T, DOWNSAMPLE_FACTOR)                  # :355
    spec = ops.stft_magnitude(x, T, 400, DOWNSAMPLE_FACTOR)                          # :326, :361-363
    assert abs(len(f0) - T) <= 1 and len(f0) >= T                                    # :385
    f0 = f0[:T].to(dev).float()
    harmonics = ops.harmonic_amplitudes(spec, f0, 49, 16000)                         # :391-404
    return matching, synth, audio, spec, f0, harmonics


def prematch_speaker(matching_list, f0_list, spec_list, harmonics_list, utterance_start_indices, fit=True):
    """Tensor-level body of per_spk_extract for ONE speaker (reference :1560-1769), all
    utterances at once: the speaker's `.half().float()` pool is matched against itself with
    every utterance's own frames masked to distance 1 (:1623-1624) in ONE fused search, then
    one f0 re-rank, one amp_ratio gather and one batched compute_weight_with_amp launch
    (one CTA per utterance).  Returns per-utterance dicts in the on-disk layout (:1750-1769)."""
    dev = matching_list.device
    offs = [int(v) for v in utterance_start_indices]
    n = offs[-1]
    assert n == len(matching_list)                                                   # :1514
    pool = ops.prepare_rows(matching_list)
    lens = torch.tensor([offs[u + 1] - offs[u] for u in range(len(offs) - 1)], device=dev)
    starts = torch.tensor(offs[:-1], device=dev, dtype=torch.int64)
    mask_lo = torch.repeat_interleave(starts, lens)
    mask_hi = torch.repeat_interleave(starts + lens, lens)
    _, nearest_nbrs = ops.knn_search(pool, pool, 32, mask_lo=mask_lo, mask_hi=mask_hi)          # :1608-1632
    f0_dev = f0_list.to(dev).float()
    prio = sort_by_f0_compatibility(f0_dev, f0_dev, nearest_nbrs)                    # :1646
    idx = prio[:, :4].contiguous()                                                   # :1655
    l1 = ops.row_l1(spec_list.to(dev))                                               # :1672-1673
    ratio = ops.amp_ratio(l1, l1, idx)                                               # :1674
    weights = compute_weight_with_amp(idx, harmonics_list.to(dev), "sum_to_1_geq", amp_ratio=ratio,
                                      utt_offsets=offs) if fit else None             # :1681
    out = []
    for u in range(len(offs) - 1):
        a, b = offs[u], offs[u + 1]
        d = {"slice": (a, b), "nearest_nbrs": nearest_nbrs[a:b], "nearest_nbrs_f0_priority": prio[a:b],
             "amp_ratio": ratio[a:b]}
        if weights is not None:
            d["harmonics_best_weight_para"] = weights[a:b]
        out.append(d)
    return out


def per_spk_extract(wavlm, device, ls_path, out_path, synth_weights, match_weights, save_pool_only=False):
    """Offline prematch, speaker by speaker — same call and same files as the reference
    (:1464-1770): `<out>/<spk>/pool.npy`, `pool_harmonics.npy` (and `pool_f0.npy`,
    `pool_spec.npy` with save_pool_only) plus one pickle per utterance
    {slice, nearest_nbrs [T,32], nearest_nbrs_f0_priority [T,32], harmonics_best_weight_para
    [T,4], amp_ratio [T,4]} that hifigan/ddsp_meldataset.py:473-486 and
    hifigan/knn_data_cnpop.py:200-207 read.  The feature producer `get_complete_spk_pool` is the
    module attribute (the reference's, unless the caller assigns another)."""
    import pickle
    from pathlib import Path

    import numpy as np
    ls_path, out_path = Path(ls_path), Path(out_path)
    audio_files = list(ls_path.glob("**/*.wav")) + list(ls_path.glob("**/*.flac"))
    spk_folders = list(set(f.parent for f in audio_files))                           # :1473
    dev = torch.device(device)
    for i, folder in enumerate(spk_folders):
        matching_pool, synth_pool, audio_synth_pool, spec_synth_pool, f0_pool, harmonics_synth_pool = \
            get_complete_spk_pool(folder, wavlm, match_weights, synth_weights, device=device)
        items = list(matching_pool)
        offs = [0]
        for item in items:
            offs.append(offs[-1] + len(matching_pool[item]))
        synth_list = torch.concat([synth_pool[k] for k in items], dim=0).half().float()          # :1510
        spec_list = torch.concat([spec_synth_pool[k] for k in items], dim=0)
        f0_list = torch.concat([f0_pool[k] for k in items], dim=0)
        harmonics_list = torch.concat([harmonics_synth_pool[k] for k in items], dim=0)
        assert offs[-1] == len(synth_list)                                           # :1514
        spk_cache_folder = out_path / folder.relative_to(ls_path)
        os.makedirs(spk_cache_folder, exist_ok=True)
        np.save(str(spk_cache_folder / "pool.npy"), synth_list.cpu().numpy())        # :1534
        np.save(str(spk_cache_folder / "pool_harmonics.npy"), harmonics_list.cpu().numpy())      # :1536
        results = None
        if save_pool_only:
            np.save(str(spk_cache_folder / "pool_f0.npy"), f0_list.cpu().numpy())    # :1595-1597
            np.save(str(spk_cache_folder / "pool_spec.npy"), spec_list.cpu().numpy())
        else:
            matching_list = torch.concat([matching_pool[k] for k in items], dim=0).to(dev).half().float()   # :1567
            results = prematch_speaker(matching_list, f0_list, spec_list, harmonics_list, offs)
        for k, item in enumerate(items):
            target = out_path / Path(item).relative_to(ls_path).with_suffix(".pt")
            os.makedirs(target.parent, exist_ok=True)
            existing = {"slice": (offs[k], offs[k + 1])}
            if os.path.isfile(target):
                with open(target, "rb") as handle:
                    existing = pickle.load(handle)
                assert existing["slice"] == (offs[k], offs[k + 1])                   # :1587
            if results is not None:
                r = results[k]
                existing["nearest_nbrs"] = r["nearest_nbrs"].cpu().numpy()           # :1754
                existing["nearest_nbrs_f0_priority"] = r["nearest_nbrs_f0_priority"].cpu().numpy()
                existing["harmonics_best_weight_para"] = r["harmonics_best_weight_para"].cpu().numpy()
                existing.pop("best_weights", None)                                   # :1761-1762
                existing["amp_ratio"] = r["amp_ratio"].cpu().numpy()
            with open(target, "wb") as handle:
                pickle.dump(existing, handle, protocol=pickle.HIGHEST_PROTOCOL)
        print(i, "/", len(spk_folders), "/".join(str(folder).split("/")[-3:]), flush=True)


def read_prematched(feat_path, device="cuda"):
    """What the training datasets read back (hifigan/ddsp_meldataset.py:473-486,
    hifigan/knn_data_cnpop.py:200-207): `pool.npy[nearest_nbrs[:, :4]].mean(1)` and the gathered
    harmonic candidates with their amp_ratio.  Returns (mel [T,D], harmonics [T,4,H], amp_ratio [T,4])."""
    import pickle
    from pathlib import Path

    import numpy as np
    feat_path = Path(feat_path)
    with open(feat_path, "rb") as handle:
        feat_dict = pickle.load(handle)
    dev = torch.device(device)
    pool = torch.from_numpy(np.load(str(feat_path.parent / "pool.npy"))).to(dev)
    mel = ops.gather_mix(pool, torch.from_numpy(feat_dict["nearest_nbrs"][:, :4].copy()).to(dev), None)
    harm_pool = torch.from_numpy(np.load(str(feat_path.parent / "pool_harmonics.npy"))).to(dev)
    hidx = torch.from_numpy(feat_dict["nearest_nbrs_f0_priority"][:, :4].copy()).to(dev)
    return mel, harm_pool[hidx], torch.from_numpy(feat_dict["amp_ratio"]).to(dev)
