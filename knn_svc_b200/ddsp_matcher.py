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
        """WavLM features (seq_len, dim) of a wav path or tensor, optional VAD trim (reference :438-518).
        One deliberate deviation: the reference aligns the VAD cut to the 320-sample hop with
        `x_front_trim[extra_cut:]` (:468, :479), which slices dim 0 — the CHANNEL axis of the [1, T]
        waveform — and leaves an empty tensor whenever the cut is not a multiple of the hop (the next
        call then crashes).  Here the SAMPLE axis is sliced, which is what the surrounding code
        (`lstrip_len += extra_cut`) means."""
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
                    trimmed = trimmed[..., self.hop_length - cut % self.hop_length:]
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
              post_opt: str = "no_post_opt", matching_f0: Tensor = None, gather: str = "all") -> Tensor:
        """kNN regression of `query_seq` onto `matching_set` (reference :522-586; its dead
        debug block :559-576 removed).  `post_opt` other than "no_post_opt" routes the
        top-k through the concatenation-smoothness stage exactly as
        match_at_inference_time does (greedy re-selection + fitted mixing weights).
        `matching_set` may be a `sharded.ShardedPool` (a pool sharded by frame over the GPUs of the
        box, one process per GPU): every rank makes the same call; with `gather="all"` (default) each
        gets all T matched rows, with `gather="slice"` only the rows it produced."""
        device = torch.device(device) if device is not None else self.device
        from .sharded import ShardedPool
        sharded_pool = matching_set if isinstance(matching_set, ShardedPool) else None
        if sharded_pool is None:
            synth_set = matching_set.to(device) if synth_set is None else synth_set.to(device)
            matching_set = matching_set.to(device)
            query_seq = query_seq.to(device)
        else:
            device = sharded_pool.device
        if target_duration is not None:
            target_samples = int(target_duration * self.sr)
            scale_factor = (target_samples / self.hop_length) / query_seq.shape[0]
            query_seq = F.interpolate(query_seq.T[None], scale_factor=scale_factor, mode="linear")[0].T
        if sharded_pool is not None:
            # a pool sharded by frame over the GPUs of the box (one process per GPU, sharded.py): every
            # rank calls match() with the same query and gets all T matched rows back, bit-identical to
            # the single-GPU result.  synth_set, if any, was given to the ShardedPool (`synth_rows`).
            if synth_set is not None:
                raise ValueError("pass the synth rows to ShardedPool(synth_rows=...), not to match()")
            if "no_post_opt" not in post_opt:
                if topk != 4:
                    raise ValueError("post_opt needs topk=4 (the reference keeps 4 candidates, ddsp_prematch_dataset.py:1246)")
                out_feats = sharded_pool.match_post_opt(query_seq, _pm.parse_post_opt(post_opt), gather=gather).feats
            else:
                out_feats = sharded_pool.match(query_seq, topk, gather=gather).feats   # host queries: upload_query
            assert gather != "all" or out_feats.shape == query_seq.shape
            if without_vocode:
                return out_feats
            f0 = None if query_f0 is None else query_f0[None, :, None].to(device)
            return self.vocode(out_feats[None], f0).squeeze()
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
        """match_at_inference_time as the reference's special_match / bulk_match call it: the
        `mix` branch forwards `post_opt` (:959, :1091), the `wavlm_only` / `no_harm_no_amp` branch
        does NOT (:970, :1112), so those checkpoints always run with "no_post_opt"."""
        from .ddsp_prematch_dataset import match_at_inference_time
        if "wavlm_only" in ckpt_type or "no_harm_no_amp" in ckpt_type:
            post_opt = "no_post_opt"
        return match_at_inference_time(Path(src_wav_file), Path(ref_wav_file), self.wavlm, match_weights=self.weighting,
                                       synth_weights=self.weighting, topk=topk, device=device,
                                       prioritize_f0=prioritize_f0, ckpt_type=ckpt_type, post_opt=post_opt, **kw)

    def _vocode_item(self, res, item, ckpt_type, device):
        """reference :961-983 / :1093-1128: which vocoder inputs each checkpoint type gets"""
        if "wavlm_only" not in ckpt_type and "no_harm_no_amp" not in ckpt_type:
            feats, harm, _, f0 = res
            return self.vocode(feats[item][None].to(device), f0[item][None, :, None], harm[item][None]).squeeze()
        feats, _, f0 = res
        if "wavlm_only_original" in ckpt_type:                      # the original kNN-VC vocoder takes no f0
            return self.vocode(feats[item][None].to(device)).squeeze()
        return self.vocode(feats[item][None].to(device), f0[item][None, :, None].to(device)).squeeze()

    def special_match(self, src_wav_file, ref_wav_file, topk: int = 4, device=None, prioritize_f0=True,
                      ckpt_type="wavlm_only", tgt_loudness_db=-16, post_opt="no_post_opt", save=True) -> Tensor:
        """One source file converted with one reference file (reference :937-1023).  Returns the
        waveform; like the reference it writes `<src>_to_<ref>_knn_<ckpt>_<post_opt>.wav` next to the
        source (the reference then calls sys.exit(), which is not reproduced)."""
        from .lib_ongaku_test import save_audio
        device = torch.device(device) if device is not None else self.device
        res = self._convert(src_wav_file, ref_wav_file, topk, device, prioritize_f0, ckpt_type, post_opt)
        key = Path(src_wav_file) if Path(src_wav_file) in res[0] else src_wav_file
        pred = self._vocode_item(res, key, ckpt_type, device)
        if save:
            src_id = os.path.basename(src_wav_file).split(".")[0]
            ref_id = os.path.basename(ref_wav_file).split(".")[0]
            out = str(Path(src_wav_file).parent) + "/" + src_id + "_to_" + ref_id + f"_knn_{ckpt_type}_{post_opt}.wav"
            print("->", out)
            save_audio(out, pred.detach().cpu().numpy(), sample_rate=16000)                 # :1017
        return pred

    def bulk_match(self, src_dataset_path, tgt_dataset_path, converted_audio_dir, topk: int = 4, device=None,
                   prioritize_f0=True, ckpt_type="mix", tgt_loudness_db=-16, required_subset_file=None,
                   post_opt="no_post_opt", duration_limit=None, cache_pools: bool = True):
        """Dataset -> dataset conversion over (source speaker, target speaker) folder pairs
        (reference :1027-1155): speaker folders are the sub-directories of the dataset roots
        (`f0_cache*` skipped), a speaker is not converted to itself when both roots are the same,
        the split file (header row skipped) lists `<utt>/<tgt_spk>` in column 2 of the rows whose
        LAST column is "0" (:1050-1054), and each prediction is written to
        `<out>/<src_spk>/<utt>/<tgt_spk>.<ext of the source file>` (:1133).  `cache_pools` keeps
        every speaker's pool (WavLM features, and for targets the HBM-resident prepared pool) across
        the pair loop instead of rebuilding both for each pair as the reference does (:1073-1112);
        results are identical.  Returns the list of files written."""
        import csv
        from .ddsp_prematch_dataset import PoolCache
        from .lib_ongaku_test import save_audio
        assert os.path.isdir(src_dataset_path) and os.path.isdir(tgt_dataset_path)
        device = torch.device(device) if device is not None else self.device
        Path(converted_audio_dir).mkdir(parents=True, exist_ok=True)
        src_spk_folders = sorted(i for i in Path(src_dataset_path).iterdir() if i.is_dir() and "f0_cache" not in i.name)
        tgt_spk_folders = sorted(i for i in Path(tgt_dataset_path).iterdir() if i.is_dir() and "f0_cache" not in i.name)
        if src_dataset_path != tgt_dataset_path:
            assert len(set(src_spk_folders).intersection(set(tgt_spk_folders))) == 0
        assert len(src_spk_folders) > 0, [f"Are you sure {src_dataset_path} is a FOLDER containing speaker folders, i.e. dataset root"]
        assert len(tgt_spk_folders) > 0, [f"Are you sure {tgt_dataset_path} is a FOLDER containing speaker folders, i.e. dataset root"]
        required = None
        if required_subset_file:
            with open(required_subset_file, "r") as fp:
                reader = csv.reader(fp, delimiter=",", quotechar='"')
                required = [row[2] for row_idx, row in enumerate(reader) if row_idx != 0 and row[-1] == "0"]
        cache = PoolCache(max_entries=len(src_spk_folders) + len(tgt_spk_folders)) if cache_pools else None
        written = []
        for i, spk_folder in enumerate(src_spk_folders):
            for j, tgt_spk_folder in enumerate(tgt_spk_folders):
                if src_dataset_path == tgt_dataset_path and i == j:      # avoid self to self
                    continue
                print(f"{spk_folder} -> {tgt_spk_folder}")
                res = self._convert(spk_folder, tgt_spk_folder, topk, device, prioritize_f0, ckpt_type, post_opt,
                                    src_dataset_path=src_dataset_path, tgt_dataset_path=tgt_dataset_path,
                                    required_subset=required, duration_limit=duration_limit, pool_cache=cache)
                for item in res[0]:
                    pred = self._vocode_item(res, item, ckpt_type, device)
                    assert len(pred.shape) == 1
                    name = os.path.basename(str(item))
                    out = os.path.join(converted_audio_dir, os.path.basename(spk_folder), name.split(".")[0],
                                       os.path.basename(tgt_spk_folder) + "." + name.split(".")[-1])
                    Path(out).parent.mkdir(parents=True, exist_ok=True)
                    save_audio(out, pred[None, :].cpu().numpy(), sample_rate=self.sr)       # :1150
                    written.append(out)
                print(f"{os.path.basename(spk_folder)}, {os.path.basename(tgt_spk_folder)} -> {converted_audio_dir}")
        return written
