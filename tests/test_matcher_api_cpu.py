"""SURVEY §8 row a12 on the CPU: our `KNeighborsVC` against the outputs of the REFERENCE's
`KNeighborsVC` (tests/golden/make_golden_prematch.py, same fake WavLM / vocoder): construction,
`get_features` (fast path and layer-weighted path), `get_matching_set`, `vocode`.  These members
keep the reference's semantics and do not touch the CUDA library, so they run without a GPU."""
import numpy as np
import torch

from knn_svc_b200 import synth
from tests.util import FakeWavLM, fake_waveform


class Cfg:
    sampling_rate = 16000


class FakeVocoder(torch.nn.Module):
    def forward(self, c, f0=None, harm=None):
        y = c.sum(-1)
        if f0 is not None:
            y = y + f0[..., 0]
        if harm is not None:
            y = y + harm.sum(-1)
        return y[:, None, :]


def _knn():
    from knn_svc_b200.ddsp_matcher import KNeighborsVC
    wavlm = FakeWavLM()
    return KNeighborsVC(wavlm, FakeVocoder(), Cfg(), device="cpu"), wavlm


def test_members_match_the_reference(golden_pm):
    knn, _ = _knn()
    assert knn.weighting.dtype == torch.float64
    assert np.array_equal(knn.weighting.numpy(), golden_pm["a12_weighting"])
    assert knn.sr == 16000 and knn.hop_length == 320 and knn.device == torch.device("cpu")


def test_get_features_and_matching_set_match_the_reference(golden_pm):
    knn, wavlm = _knn()
    wav = fake_waveform()
    fast = knn.get_features(wav, None, 0)
    assert wavlm.calls[-1][1:] == (6, False)                            # layer-6 fast path, as the reference takes
    assert np.array_equal(fast.numpy(), golden_pm["a12_feats_fast"])
    w2 = torch.linspace(0.0, 1.0, 25, dtype=torch.float64)[:, None]
    slow = knn.get_features(wav, w2, 0)
    assert wavlm.calls[-1][1:] == (24, True)                            # all layers, weighted sum
    assert slow.dtype == torch.float64
    assert np.abs(slow.numpy() - golden_pm["a12_feats_weighted"]).max() < 1e-12
    ms = knn.get_matching_set([wav, wav[:16000]], None, 0)
    assert ms.device.type == "cpu" and np.array_equal(ms.numpy(), golden_pm["a12_matching_set"])
    feats, audio = knn.get_features(wav, None, 0, return_audio=True)
    assert audio.shape == (1, len(wav)) and torch.equal(feats, fast)


def test_vocode_matches_the_reference(golden_pm):
    knn, _ = _knn()
    c = torch.from_numpy(synth.randn_frames(12, 16, seed=77))[None]
    f0v = torch.from_numpy(synth.f0_track(12, seed=78))[None, :, None]
    hv = torch.from_numpy(synth.harmonics_pool(12, seed=79))[None]
    assert np.array_equal(knn.vocode(c).numpy(), golden_pm["a12_vocode_plain"])
    assert np.array_equal(knn.vocode(c, f0v).numpy(), golden_pm["a12_vocode_f0"])
    assert np.array_equal(knn.vocode(c, f0v, hv).numpy(), golden_pm["a12_vocode_mix"])


def test_get_features_vad_trim_keeps_the_sample_axis():
    """the reference slices dim 0 (channels) when it aligns the VAD cut to the hop
    (ddsp_matcher.py:468,479), which empties the waveform whenever the cut is not a multiple of 320.
    Ours slices the sample axis: features come back, and each side loses the VAD's own cut rounded UP
    to a whole number of hops."""
    import torchaudio.transforms as T
    knn, wavlm = _knn()
    g = torch.Generator().manual_seed(9)
    t = torch.arange(32000) / 16000.0
    voiced = 0.5 * torch.sin(2 * np.pi * 220 * t) * (1 + 0.5 * torch.sin(2 * np.pi * 3 * t)) \
        + 0.2 * torch.randn(32000, generator=g)
    wav = torch.cat([1e-4 * torch.randn(7777, generator=g), voiced, 1e-4 * torch.randn(3000, generator=g)])
    vad = T.Vad(sample_rate=16000, trigger_level=7)
    front = len(wav) - vad(wav[None]).shape[-1]
    assert front > 0 and front % 320 != 0, "fixture: the raw VAD cut must not be hop-aligned"
    feats, audio = knn.get_features(wav, None, vad_trigger_level=7, return_audio=True)
    assert audio.dim() == 2 and audio.shape[0] == 1 and audio.shape[1] > 16000
    cut = len(wav) - audio.shape[1]
    assert cut >= front + (320 - front % 320)
    assert torch.equal(audio[0, :100], wav[front + (320 - front % 320):][:100])     # front edge: hop-aligned cut
    assert feats.shape[0] == (audio.shape[1] - 400) // 320 + 1
