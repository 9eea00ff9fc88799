"""Golden fixtures for the SURVEY §8(f) rows (offline prematch, compute_weight_with_amp,
pool-builder tensor ops), produced by running THE REFERENCE ITSELF (read-only import from
/root/reference) in the build container:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_prematch.py

* pool builder: the reference's own `get_complete_spk_pool` (ddsp_prematch_dataset.py:301-423)
  runs on a 2-second crop of `sample_content/Tiken_lead_07…wav` with its real f0 track; only
  the two things absent from the container are stubbed — `torchaudio.load` (needs torchcodec)
  is replaced by a stdlib `wave` reader and `get_full_wavlm_features` (needs the WavLM
  checkpoint) returns seeded synthetic layer features.  The crop (int16) is stored so the
  tests are self-contained.
* offline prematch: the reference's own `per_spk_extract` (:1464-1770) runs end to end on a
  temporary folder with `get_complete_spk_pool` stubbed to return seeded synthetic pools; the
  files it writes (pool.npy, pool_harmonics.npy, per-utterance pickles) are the fixtures.
* compute_weight_with_amp (:684-803) is also run directly, fp32 and fp64.
"""
import contextlib
import io
import os
import pickle
import re
import sys
import tempfile
import wave
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True

import torchaudio                          # noqa: E402
import ddsp_prematch_dataset as ref_pm     # noqa: E402  (reference)
from knn_svc_b200 import synth             # noqa: E402

torch.manual_seed(0)
torch.set_num_threads(8)

SC = Path("/root/reference/sample_content")
WAV = SC / "Tiken_lead_07_resampled_16000_cut.wav"
F0 = SC / "Tiken_lead_07_resampled_16000_cut_f0.npy"


def quiet(fn, *a, **k):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        out = fn(*a, **k)
    return out, buf.getvalue()


def wave_load(path, *a, **k):
    """stdlib replacement for torchaudio.load (torchcodec is not installed): int16 PCM -> [C, N] float32 in [-1, 1)"""
    with wave.open(str(path), "rb") as w:
        assert w.getsampwidth() == 2
        sr, ch, n = w.getframerate(), w.getnchannels(), w.getnframes()
        pcm = np.frombuffer(w.readframes(n), dtype="<i2").reshape(n, ch).T
    return torch.from_numpy(pcm.astype(np.float32) / 32768.0), sr


def prematch_inputs():
    """seeded synthetic speaker: 3 utterances, shared by the golden script and the tests"""
    lens = [60, 90, 50]
    feats = synth.ar1_frames(sum(lens), seed=61, reset_every=70)
    spec = np.abs(synth.randn_frames(sum(lens), 200, seed=62)).astype(np.float32) + 0.05
    harm = synth.harmonics_pool(sum(lens), seed=63)
    f0 = np.load(F0).astype(np.float32)[400:400 + sum(lens)].copy()
    return lens, feats, spec, harm, f0


def main():
    out = {}
    torchaudio.load = wave_load

    # ------------------------------------------------------------------ pool builder on real audio
    T = 100
    n_samples = 320 * (T - 1) + 400                     # WavLM conv stack: floor((N-400)/320)+1 = T frames
    first = 320 * 1000                                  # crop starts at frame 1000 of the 60 s file
    x_full, sr = wave_load(WAV)
    assert sr == 16000
    pcm = np.round(x_full[0, first:first + n_samples].numpy() * 32768.0).astype(np.int16)
    f0_crop = np.load(F0).astype(np.float32)[1000:1000 + T + 1].copy()
    f0_crop[40:46] = 0.0                                # make sure the unvoiced branch (:401-402) is exercised
    L, D = 25, 64
    layer_feats = synth.randn_frames(L * T, D, seed=71).reshape(L, T, D)
    rs = np.random.RandomState(72)
    match_w = rs.rand(L, 1)
    match_w /= match_w.sum()
    synth_w = np.zeros((L, 1)); synth_w[6] = 1.0        # knnvc_utils.generate_matrix_from_index(6): float64 one-hot
    ref_pm.get_full_wavlm_features = lambda x, sr, wavlm, device: torch.from_numpy(layer_feats)
    with tempfile.TemporaryDirectory() as td:
        wav = Path(td) / "crop.wav"
        with wave.open(str(wav), "wb") as w:
            w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000); w.writeframes(pcm.tobytes())
        np.save(Path(td) / "crop_f0.npy", f0_crop)
        res, _ = quiet(ref_pm.get_complete_spk_pool, wav, None, torch.from_numpy(match_w), torch.from_numpy(synth_w),
                       device="cpu")
    matching_pool, synth_pool, audio_pool, spec_pool, f0_pool, harm_pool = res
    key = next(iter(matching_pool))
    out["pb_pcm"] = pcm
    out["pb_f0"] = f0_crop
    out["pb_match_w"], out["pb_synth_w"] = match_w, synth_w
    out["pb_matching"] = matching_pool[key].numpy()                 # float64 [T, D]
    out["pb_synth"] = synth_pool[key].numpy()
    out["pb_spec"] = spec_pool[key].numpy()                         # float32 [T, 200]
    out["pb_harmonics"] = harm_pool[key].numpy()                    # float32 [T, 49]
    assert out["pb_spec"].shape == (T, 200) and out["pb_harmonics"].shape == (T, 49)
    # the x8 interpolation on its own (one row), ddsp_prematch_dataset.py:395
    out["pb_interp_row0"] = torch.nn.functional.interpolate(spec_pool[key][None, :1], scale_factor=8,
                                                            mode="linear").squeeze(0).numpy()

    # ------------------------------------------------------------------ compute_weight_with_amp directly
    T6 = 80
    p6 = synth.harmonics_pool(400, seed=33)
    rs = np.random.RandomState(34)
    idx6 = torch.from_numpy(rs.randint(0, 400, size=(T6, 4)).astype(np.int64))
    amp6 = (0.5 + rs.rand(T6, 4) * 1.5).astype(np.float32)
    out["k6amp_idx"], out["k6amp_amp"] = idx6.numpy(), amp6
    for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        w, log = quiet(ref_pm.compute_weight_with_amp, idx6.clone(), torch.from_numpy(p6).to(dt), "sum_to_1_geq",
                       amp_ratio=torch.from_numpy(amp6).to(dt))
        its = [int(m) for m in re.findall(r"(?:^|\r|\n)(\d+) ", log)]
        out[f"k6amp_w_{tag}"] = w.detach().numpy()
        out[f"k6amp_last_t_{tag}"] = np.int64(max(its))

    # ------------------------------------------------------------------ per_spk_extract end to end
    lens, feats, spec, harm, f0 = prematch_inputs()
    out["pm_f0"] = f0
    with tempfile.TemporaryDirectory() as td:
        ls, outp = Path(td) / "in", Path(td) / "out"
        spk = ls / "spk0"
        spk.mkdir(parents=True)
        names = [spk / f"utt{i}.wav" for i in range(len(lens))]
        for n in names:
            n.touch()

        def fake_pool(path, *a, **k):
            pools = [dict() for _ in range(6)]
            o = 0
            for n, ln in zip(names, lens):
                sl = slice(o, o + ln)
                vals = (torch.from_numpy(feats[sl]), torch.from_numpy(feats[sl]), torch.zeros(ln, 320),
                        torch.from_numpy(spec[sl]), torch.from_numpy(f0[sl]), torch.from_numpy(harm[sl]))
                for d, v in zip(pools, vals):
                    d[str(n)] = v
                o += ln
            return tuple(pools)

        ref_pm.get_complete_spk_pool = fake_pool
        quiet(ref_pm.per_spk_extract, None, "cpu", ls, outp, None, None)
        out["pm_pool"] = np.load(outp / "spk0" / "pool.npy")[:, ::16]
        out["pm_pool_harmonics"] = np.load(outp / "spk0" / "pool_harmonics.npy")
        for i in range(len(lens)):
            with open(outp / "spk0" / f"utt{i}.pt", "rb") as fh:
                d = pickle.load(fh)
            out[f"pm_u{i}_slice"] = np.asarray(d["slice"], dtype=np.int64)
            assert sorted(d) == ["amp_ratio", "harmonics_best_weight_para", "nearest_nbrs",
                                 "nearest_nbrs_f0_priority", "slice"], sorted(d)
            for k2 in ("nearest_nbrs", "nearest_nbrs_f0_priority", "harmonics_best_weight_para", "amp_ratio"):
                out[f"pm_u{i}_{k2}"] = np.asarray(d[k2])

    # ------------------------------------------------------------------ a12: the reference's KNeighborsVC itself
    import ddsp_matcher as ref_matcher          # (reference)
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

    knn = ref_matcher.KNeighborsVC(FakeWavLM(), FakeVocoder(), Cfg(), device="cpu")
    wav = fake_waveform()
    out["a12_weighting"] = knn.weighting.numpy()
    f_fast, _ = quiet(knn.get_features, wav, None, 0)
    out["a12_feats_fast"] = f_fast.numpy()
    w2 = torch.linspace(0.0, 1.0, 25, dtype=torch.float64)[:, None]
    f_slow, _ = quiet(knn.get_features, wav, w2, 0)
    out["a12_feats_weighted"] = f_slow.numpy()
    ms, _ = quiet(knn.get_matching_set, [wav, wav[:16000]], None, 0)
    out["a12_matching_set"] = ms.numpy()
    c = torch.from_numpy(synth.randn_frames(12, 16, seed=77))[None]
    f0v = torch.from_numpy(synth.f0_track(12, seed=78))[None, :, None]
    hv = torch.from_numpy(synth.harmonics_pool(12, seed=79))[None]
    out["a12_vocode_plain"] = knn.vocode(c).numpy()
    out["a12_vocode_f0"] = knn.vocode(c, f0v).numpy()
    out["a12_vocode_mix"] = knn.vocode(c, f0v, hv).numpy()

    np.savez_compressed(HERE / "prematch_outputs.npz", **out)
    print({k: (v.shape, str(v.dtype)) for k, v in out.items()})
    print("bytes", os.path.getsize(HERE / "prematch_outputs.npz"))


if __name__ == "__main__":
    main()
