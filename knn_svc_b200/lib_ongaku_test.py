"""Drop-in replacements for the matcher functions of the reference's
`lib_ongaku_test.py`, backed by the CUDA library.  Same names, argument order
and return conventions; CUDA tensors only (no CPU fallback).

    from knn_svc_b200.lib_ongaku_test import fast_cosine_dist, knn_with_concat_cost
"""
from __future__ import annotations

import torch

from . import ops


def fast_cosine_dist(source_feats_collection, matching_pool, increment: int = 20):
    """[T, Np] cosine distances — reference lib_ongaku_test.py:148-175.

    `increment` (the reference's 20-row chunk) is accepted and ignored: chunking
    only bounded the reference's temporaries.  NaN distances (a zero-norm row)
    make the reference print "containing nan" and exit (:166-169); here they
    raise ValueError before any launch result is used.  The fused kNN
    (`knn_cosine_similarity`, `ops.knn_search`) never builds this matrix; this
    function exists for callers that want the matrix itself.
    """
    q = source_feats_collection
    p = matching_pool
    if q.device != p.device:
        p = p.to(q.device)
    out = ops.cosine_dist(q, p)
    if torch.isnan(out).any():
        raise ValueError("containing nan")
    return out.to(q.dtype) if q.dtype in (torch.float64, torch.float16, torch.bfloat16) else out


def knn_cosine_similarity(src_elements, tgt_elements, retain_mask=None, topk: int = 32):
    """(indices, values) of the `topk` nearest pool rows — reference
    lib_ongaku_test.py:182-199, which rounds both sides through fp16 first
    (`.half().float()`).  `retain_mask` adds (1 - mask) to the distances, which
    needs the full matrix, so that variant goes through `fast_cosine_dist`."""
    s = src_elements.half().float()
    t = tgt_elements.half().float()
    if retain_mask is not None:
        d = fast_cosine_dist(s, t)
        if retain_mask.shape != d.shape:
            raise AssertionError("retain_mask shape mismatch")
        best = (d + (1 - retain_mask.to(d))).topk(k=topk, dim=-1, largest=False)
        return best.indices, best.values
    dist, idx = ops.knn_search(ops.prepare_rows(s), ops.prepare_rows(t), topk)
    return idx, dist


def knn_with_concat_cost(target_feature_indices, src_elements, tgt_elements, shifted_src_f0=None, tgt_f0=None,
                         concat_weight: float = 0.2):
    """Greedy concatenation-cost re-selection — reference lib_ongaku_test.py:270-369.
    Returns [T, 4] int64 indices on the device of `src_elements`."""
    if len(target_feature_indices) != len(src_elements):
        raise AssertionError("indices and src_elements must have the same number of frames")
    if shifted_src_f0 is not None and tgt_f0 is None:
        raise AssertionError("tgt_f0 is required with shifted_src_f0")
    return ops.concat_cost_reselect(target_feature_indices, src_elements, tgt_elements,
                                    shifted_src_f0 if shifted_src_f0 is not None else None,
                                    tgt_f0 if shifted_src_f0 is not None else None,
                                    concat_weight=concat_weight)


def save_audio(filename, waveform, sample_rate):
    """reference lib_ongaku_test.py:89-145: float audio in [-1, 1] (rescaled by its peak when it
    exceeds 1) to 32-bit PCM.  `.wav` is written as PCM_32 with the standard library (the reference
    uses soundfile, same bytes on disk: RIFF/WAVE, 32-bit signed little-endian); `.mp3` / `.flac`
    need pydub + ffmpeg as in the reference and raise if pydub is missing."""
    import numpy as np
    if isinstance(waveform, torch.Tensor):
        waveform = waveform.detach().cpu().numpy()
    if waveform.dtype == np.float32 or waveform.dtype == np.float64:
        peak = np.max(np.abs(waveform)) if waveform.size else 0.0
        if peak > 1:
            waveform = waveform / peak
        waveform = (waveform * (2 ** 31 - 1)).astype(np.int32)
    else:
        assert waveform.dtype == np.int32
    if waveform.ndim == 2 and waveform.shape[0] in {1, 2}:
        waveform = waveform.T                                    # [samples, channels]
    channels = 1 if waveform.ndim == 1 else waveform.shape[1]
    if filename.endswith(".wav"):
        import wave
        with wave.open(filename, "wb") as w:
            w.setnchannels(channels)
            w.setsampwidth(4)
            w.setframerate(int(sample_rate))
            w.writeframes(np.ascontiguousarray(waveform).astype("<i4").tobytes())
        return
    assert filename.split(".")[-1] in {"mp3", "flac"}
    from pydub import AudioSegment                                # as the reference (:137-145)
    AudioSegment(waveform.tobytes(), frame_rate=sample_rate, sample_width=4, channels=channels).export(
        filename, format=filename.split(".")[-1], bitrate="320k")
