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
    pad = torch.zeros((U, L), dtype=f0_list[0].dtype if U else torch.float32, device=dev)
    for u, f in enumerate(f0_list):
        pad[u, :lens[u]] = f.to(dev)
    voiced = pad != 0
    logf = torch.log(torch.where(voiced, pad, torch.ones_like(pad)))
    med = _lower_median_rows(logf, voiced)                      # NaN-free even for all-unvoiced rows
    shifted = torch.where(voiced, torch.exp(logf + matching_f0_median.to(pad.dtype) - med[:, None]), pad)
    keep = torch.arange(L, device=dev)[None, :] < torch.tensor(lens, device=dev)[:, None]
    return shifted[keep]


def parse_post_opt(post_opt: str) -> float:
    """reference :1273-1279"""
    try:
        return float(post_opt.split("_")[-1])
    except ValueError:
        return 0.3 if post_opt.split("_")[-1] == "extra" else -1


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
    dev = pool.device
    lens = [int(q.shape[0]) for q in query_seqs]
    if not lens:
        return []
    offs = [0]
    for n in lens:
        offs.append(offs[-1] + n)
    query = ops.prepare_rows(torch.concat([q.to(dev) for q in query_seqs], dim=0))
    _, nearest_nbrs = ops.knn_search(query, pool.matching, 32)               # :1196-1206
    shifted_f0 = shift_query_f0_batched(query_f0s, pool.log_f0_median)       # :1224-1233
    concat_weight = parse_post_opt(post_opt)
    fit = "no_post_opt" not in post_opt
    idx_w = nearest_nbrs[:, :4].contiguous()                                 # :1246
    if concat_weight != -1:
        idx_w = ops.concat_cost_reselect(idx_w, query.rows, pool.matching.rows, concat_weight=concat_weight,
                                         utt_offsets=offs)                   # :1295
    w = compute_wavlm_weight(idx_w, pool.synth, "sum_to_1_geq", utt_offsets=offs) if fit else None   # :1357 / :1361
    out_feats = ops.gather_mix(pool.synth, idx_w, w)                         # :1358 / :1364
    prio = sort_by_f0_compatibility(shifted_f0, pool.f0_dev, nearest_nbrs)   # :1377
    idx_h = prio[:, :4].contiguous()                                         # :1398
    if concat_weight != -1:
        idx_h = ops.concat_cost_reselect(idx_h, query.rows, pool.matching.rows, shifted_f0, pool.f0_dev,
                                         concat_weight=concat_weight, utt_offsets=offs)      # :1414
    harm = None
    if "wavlm_only" not in ckpt_type and "no_harm_no_amp" not in ckpt_type:
        hw = compute_extended_weight(idx_h, pool.harmonics, "sum_to_1_geq", [1], utt_offsets=offs) if fit else None
        harm = ops.gather_mix(pool.harmonics, idx_h, hw)                     # :1444 / :1446
    results = []
    for u in range(len(lens)):
        a, b = offs[u], offs[u + 1]
        r = {"out_feats": out_feats[a:b], "shifted_f0": shifted_f0[a:b].to(query_f0s[u].device),
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


def match_at_inference_time(src_wav_file, ref_wav_file, wavlm, match_weights, synth_weights, topk: int = 4,
                            device="cuda", prioritize_f0=False, ckpt_type="wavlm_only", src_dataset_path=None,
                            tgt_dataset_path=None, cache_dir=None, required_subset=None, post_opt="no_post_opt",
                            duration_limit=None):
    """Same call, same returns as the reference (ddsp_prematch_dataset.py:1074-1459):
    dicts keyed by query file of matched features [T,D] fp32, (mix only) mixed
    harmonics [T,49], a dict of None, and the shifted f0 [T]."""
    if src_dataset_path is None:
        assert os.path.isfile(src_wav_file)
    query_pool, _, _, query_spec_pool, query_f0_pool, _ = get_complete_spk_pool(
        src_wav_file, wavlm, match_weights, synth_weights, device=device)
    if tgt_dataset_path is None:
        assert os.path.isfile(ref_wav_file)
    matching_pool, synth_pool, audio_synth_pool, spec_synth_pool, f0_pool, harmonics_synth_pool = \
        get_complete_spk_pool(ref_wav_file, wavlm, match_weights, synth_weights, device=device,
                              duration_limit=duration_limit)
    keys = list(matching_pool)
    pool = MatchingPool(torch.concat([matching_pool[k] for k in keys], dim=0),
                        torch.concat([synth_pool[k] for k in keys], dim=0),
                        torch.concat([f0_pool[k] for k in keys], dim=0),
                        torch.concat([harmonics_synth_pool[k] for k in keys], dim=0), device)
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
