"""`KNeighborsVC` with the reference's interface (ddsp_matcher.py:303-1155), the
matching internals re-pointed at the CUDA library.

    from knn_svc_b200.ddsp_matcher import KNeighborsVC, fast_cosine_dist

WavLM (`get_features`) and the HiFi-GAN/DDSP vocoder (`vocode`) are the models
the caller passes in, exactly as in the reference; only the matcher changes.
"""
from __future__ import annotations

import os
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from . import ops
from .ddsp_prematch_dataset import MatchingPool, match_utterance
from . import ddsp_prematch_dataset as _pm

SPEAKER_INFORMATION_LAYER = 6


def _speaker_weights(n_layers: int = 25) -> np.ndarray:
    """float64 one-hot over WavLM layers (reference knnvc_utils.py:3-6, ddsp_matcher.py:88-89)."""
    w = np.zeros(n_layers, dtype=float)
    w[SPEAKER_INFORMATION_LAYER] = 1
    return w


SPEAKER_INFORMATION_WEIGHTS = _speaker_weights()


def fast_cosine_dist(source_feats: Tensor, matching_pool: Tensor, device: str = "cuda") -> Tensor:
    """Unchunked twin of lib_ongaku_test.fast_cosine_dist — reference ddsp_matcher.py:213-221."""
    dev = torch.device(device)
    return ops.cosine_dist(source_feats.to(dev), matching_pool.to(dev))


class KNeighborsVC(nn.Module):

    def __init__(self, wavlm, hifigan, hifigan_cfg, device="cuda") -> None:
        """kNN-VC matcher (reference ddsp_matcher.py:305-327): same members."""
        super().__init__()
        self.weighting = torch.tensor(SPEAKER_INFORMATION_WEIGHTS, device=device)[:, None]
        self.hifigan = hifigan.eval() if hifigan is not None else None
        self.h = hifigan_cfg
        self.wavlm = wavlm.eval() if wavlm is not None else None
        self.device = torch.device(device)
        self.sr = getattr(hifigan_cfg, "sampling_rate", 16000) if hifigan_cfg is not None else 16000
        self.hop_length = 320

    # ------------------------------------------------------------------ feature producers (unchanged semantics)
    def get_matching_set(self, wavs, weights=None, vad_trigger_level=7) -> Tensor:
        """Concatenated WavLM features of `wavs` (reference :331-343)."""
        feats = [self.get_features(p, weights=self.weighting if weights is None else weights,
                                   vad_trigger_level=vad_trigger_level) for p in wavs]
        return torch.concat(feats, dim=0).cpu()

    @torch.inference_mode()
    def vocode(self, c: Tensor, f0=None, harmonics_out_feats_weighted=None) -> Tensor:
        """Vocode features with the caller's vocoder (reference :375-394); c is (bs, seq_len, c_dim)."""
        if f0 is not None:
            if harmonics_out_feats_weighted is not None:
                y = self.hifigan(c, f0.to(c), harmonics_out_feats_weighted.to(c))
            else:
                y = self.hifigan(c, f0)
        else:
            y = self.hifigan(c)
        return y.squeeze(1)

    @torch.inference_mode()
    def get_features(self, path, weights=None, vad_trigger_level=0, return_audio=False):
        """WavLM features (seq_len, dim) of a wav path or tensor, optional VAD trim (reference :438-518)."""
        import torchaudio
        import torchaudio.transforms as T
        if weights is None:
            weights = self.weighting
        if type(path) in [str, Path] or isinstance(path, os.PathLike):
            x, sr = torchaudio.load(path, normalize=True)
            if x.dim() == 2:
                x = x[0][None, :]
        else:
            x, sr = path, self.sr
            if x.dim() == 1:
                x = x[None]
        if sr != self.sr:
            x = torchaudio.functional.resample(x, orig_freq=sr, new_freq=self.sr)
            sr = self.sr
        if vad_trigger_level > 1e-3:
            vad = T.Vad(sample_rate=sr, trigger_level=vad_trigger_level)

            def front_trim(w):
                trimmed = vad(w)
                cut = w.shape[-1] - trimmed.shape[-1]
                if cut % self.hop_length != 0:
                    trimmed = trimmed[self.hop_length - cut % self.hop_length:]
                return trimmed
            x = torch.flip(front_trim(torch.flip(front_trim(x), (-1,))), (-1,))
        wav = x.to(self.device)
        if torch.allclose(weights, self.weighting):
            feats = self.wavlm.extract_features(wav, output_layer=SPEAKER_INFORMATION_LAYER,
                                                ret_layer_results=False)[0].squeeze(0)
        else:
            _, layer_results = self.wavlm.extract_features(wav, output_layer=self.wavlm.cfg.encoder_layers,
                                                           ret_layer_results=True)[0]
            feats = torch.cat([v.transpose(0, 1) for v, _ in layer_results], dim=0)
            feats = (feats * weights[:, None]).sum(dim=0)
        return (feats, wav) if return_audio else feats

    # ------------------------------------------------------------------ matcher (the hot path)
    @torch.inference_mode()
    def match(self, query_seq: Tensor, matching_set: Tensor, query_f0: Tensor = None, synth_set: Tensor = None,
              topk: int = 4, tgt_loudness_db=-16, target_duration=None, device=None, without_vocode=False,
              post_opt: str = "no_post_opt", matching_f0: Tensor = None) -> Tensor:
        """kNN regression of `query_seq` onto `matching_set` (reference :522-586; its dead
        debug block :559-576 removed).  `post_opt` other than "no_post_opt" routes the
        top-k through the concatenation-smoothness stage exactly as
        match_at_inference_time does (greedy re-selection + fitted mixing weights)."""
        device = torch.device(device) if device is not None else self.device
        synth_set = matching_set.to(device) if synth_set is None else synth_set.to(device)
        matching_set = matching_set.to(device)
        query_seq = query_seq.to(device)
        if target_duration is not None:
            target_samples = int(target_duration * self.sr)
            scale_factor = (target_samples / self.hop_length) / query_seq.shape[0]
            query_seq = F.interpolate(query_seq.T[None], scale_factor=scale_factor, mode="linear")[0].T
        q = ops.prepare_rows(query_seq)
        p = ops.prepare_rows(matching_set)
        _, idx = ops.knn_search(q, p, topk)                                   # :550-554
        synth = p.rows if synth_set.data_ptr() == matching_set.data_ptr() else synth_set
        if "no_post_opt" in post_opt:
            out_feats = ops.gather_mix(synth, idx, None)                      # synth_set[best.indices].mean(1), :578
        else:
            cw = _pm.parse_post_opt(post_opt)
            if topk != 4:
                raise ValueError("post_opt needs topk=4 (the reference keeps 4 candidates, ddsp_prematch_dataset.py:1246)")
            if cw != -1:
                idx = ops.concat_cost_reselect(idx, q.rows, p.rows, concat_weight=cw)
            out_feats = ops.gather_mix(synth, idx, _pm.compute_wavlm_weight(idx, synth))
        assert out_feats.shape == query_seq.shape
        if without_vocode:
            return out_feats
        f0 = None if query_f0 is None else query_f0[None, :, None].to(device)
        return self.vocode(out_feats[None], f0).squeeze()

    def _convert(self, src_wav_file, ref_wav_file, topk, device, prioritize_f0, ckpt_type, post_opt, **kw):
        from .ddsp_prematch_dataset import match_at_inference_time
        res = match_at_inference_time(Path(src_wav_file), Path(ref_wav_file), self.wavlm, match_weights=self.weighting,
                                      synth_weights=self.weighting, topk=topk, device=device,
                                      prioritize_f0=prioritize_f0, ckpt_type=ckpt_type, post_opt=post_opt, **kw)
        return res

    def special_match(self, src_wav_file, ref_wav_file, topk: int = 4, device=None, prioritize_f0=True,
                      ckpt_type="wavlm_only", tgt_loudness_db=-16, post_opt="no_post_opt", save=True) -> Tensor:
        """One source file converted with one reference file (reference :937-1023).  Returns the
        waveform; like the reference it writes `<src>_to_<ref>_knn_<ckpt>_<post_opt>.wav` next to the
        source (the reference then calls sys.exit(), which is not reproduced)."""
        device = torch.device(device) if device is not None else self.device
        res = self._convert(src_wav_file, ref_wav_file, topk, device, prioritize_f0, ckpt_type, post_opt)
        key = Path(src_wav_file)
        pick = lambda d: d[key] if key in d else d[src_wav_file]  # noqa: E731
        if "wavlm_only" not in ckpt_type and "no_harm_no_amp" not in ckpt_type:
            feats, harm, _, f0 = res
            pred = self.vocode(pick(feats)[None].to(device), pick(f0)[None, :, None], pick(harm)[None]).squeeze()
        else:
            feats, _, f0 = res
            pred = self.vocode(pick(feats)[None].to(device), pick(f0)[None, :, None].to(device)).squeeze()
        if save:
            import torchaudio
            src_id = os.path.basename(src_wav_file).split(".")[0]
            ref_id = os.path.basename(ref_wav_file).split(".")[0]
            out = str(Path(src_wav_file).parent) + "/" + src_id + "_to_" + ref_id + f"_knn_{ckpt_type}_{post_opt}.wav"
            torchaudio.save(out, pred.detach().cpu()[None].float(), 16000)
        return pred

    def bulk_match(self, src_dataset_path, tgt_dataset_path, converted_audio_dir, topk: int = 4, device=None,
                   prioritize_f0=True, ckpt_type="mix", tgt_loudness_db=-16, required_subset_file=None,
                   post_opt="no_post_opt", duration_limit=None, cache_pools: bool = True):
        """Dataset -> dataset conversion over (source speaker, target speaker) folder pairs
        (reference :1027-1150).  The split file lists `<utt>/<tgt_spk>` pairs in column 2 of
        rows labelled "0".  `cache_pools` keeps every speaker's pool (WavLM features, and for
        targets the HBM-resident prepared pool) across the pair loop instead of rebuilding both
        for each pair as the reference does (:1073-1112); results are identical."""
        import csv
        import torchaudio
        from .ddsp_prematch_dataset import PoolCache
        device = torch.device(device) if device is not None else self.device
        cache = PoolCache() if cache_pools else None
        required = None
        if required_subset_file is not None:
            with open(required_subset_file) as fh:
                required = {row[2] for row in csv.reader(fh) if len(row) > 2 and row[0] == "0"}
        written = []
        for src_spk in sorted(os.listdir(src_dataset_path)):
            for tgt_spk in sorted(os.listdir(tgt_dataset_path)):
                s_dir, t_dir = os.path.join(src_dataset_path, src_spk), os.path.join(tgt_dataset_path, tgt_spk)
                if not (os.path.isdir(s_dir) and os.path.isdir(t_dir)):
                    continue
                res = self._convert(s_dir, t_dir, topk, device, prioritize_f0, ckpt_type, post_opt,
                                    src_dataset_path=src_dataset_path, tgt_dataset_path=tgt_dataset_path,
                                    required_subset=required, duration_limit=duration_limit, pool_cache=cache)
                feats, f0 = res[0], res[-1]
                harm = res[1] if len(res) == 4 else None
                for item in feats:
                    h = None if harm is None else harm[item][None]
                    pred = self.vocode(feats[item][None].to(device), f0[item][None, :, None].to(device), h).squeeze()
                    utt = os.path.basename(str(item)).split(".")[0]
                    out_dir = os.path.join(converted_audio_dir, src_spk, utt)
                    os.makedirs(out_dir, exist_ok=True)
                    out = os.path.join(out_dir, tgt_spk + ".wav")
                    torchaudio.save(out, pred.detach().cpu()[None].float(), 16000)
                    written.append(out)
        return written
