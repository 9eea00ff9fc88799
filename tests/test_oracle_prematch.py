"""Pin the oracle's §8(f) functions (masked kNN, compute_weight_with_amp, pool-builder
tensor ops, the offline prematch body) against outputs of the reference itself
(tests/golden/prematch_outputs.npz, made by tests/golden/make_golden_prematch.py).  CPU only."""
import numpy as np

from knn_svc_b200 import synth
from oracle import matcher_oracle as orc
from tests.util import GAP, pool_builder_inputs, positions_untied, prematch_inputs, set_rows


def test_layer_mix_matches_reference(golden_pm):
    _, feats, _ = pool_builder_inputs(golden_pm)
    for w, key in ((golden_pm["pb_match_w"], "pb_matching"), (golden_pm["pb_synth_w"], "pb_synth")):
        got = orc.layer_mix(feats, w)
        assert np.abs(got - golden_pm[key]).max() < 1e-12
    assert np.array_equal(orc.layer_mix(feats, golden_pm["pb_synth_w"]), feats[6].astype(np.float64))


def test_stft_magnitude_matches_reference(golden_pm):
    x, _, _ = pool_builder_inputs(golden_pm)
    ref = golden_pm["pb_spec"]
    got = orc.stft_magnitude(x, n_frames=ref.shape[0])
    assert got.shape == ref.shape == (100, 200)
    assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()


def test_interp_and_harmonic_amplitudes_match_reference(golden_pm):
    _, _, f0 = pool_builder_inputs(golden_pm)
    spec = golden_pm["pb_spec"]
    up = orc.interp_linear(spec[:1], 8)
    assert np.array_equal(up, golden_pm["pb_interp_row0"])                    # bit-exact fp32
    got = orc.harmonic_amplitudes(spec, f0)
    assert (f0 == 0).any() and (f0 != 0).any()
    assert np.array_equal(got, golden_pm["pb_harmonics"])                     # bit-exact: gather of fp32 values


def test_weight_fit_with_amp_matches_reference(golden_pm):
    pool = synth.harmonics_pool(400, seed=33)
    idx, amp = golden_pm["k6amp_idx"], golden_pm["k6amp_amp"]
    w, info = orc.compute_weight_with_amp(idx, pool, amp_ratio=amp, return_info=True)
    rows = [r * amp.astype(np.float64)[..., None] for r in orc._neighbour_rows(idx, np.asarray(pool, np.float64))]
    for tag in ("f64", "f32"):
        ref_w = golden_pm[f"k6amp_w_{tag}"]
        assert info["stop_iter"] == int(golden_pm[f"k6amp_last_t_{tag}"]) + 1
        l_ref = orc.smoothness_loss(ref_w.astype(np.float64), rows, 1000.0)
        l_got = orc.smoothness_loss(w.astype(np.float64), rows, 1000.0)
        assert abs(l_ref - l_got) <= 1e-5 * abs(l_ref) + 1e-7, (l_ref, l_got)
        assert np.abs(w - ref_w).max() < 5e-2                                 # ill-conditioned trajectory
    assert np.allclose(w.sum(1), 1, atol=1e-6)


def test_prematch_utterances_match_reference(golden_pm):
    lens, feats, spec, harm, f0 = prematch_inputs(golden_pm)
    pool = orc.half_round(feats)                                              # :1510 / :1567
    assert np.array_equal(pool[:, ::16], golden_pm["pm_pool"])
    assert np.array_equal(harm, golden_pm["pm_pool_harmonics"])
    start = 0
    for u, ln in enumerate(lens):
        end = start + ln
        assert tuple(golden_pm[f"pm_u{u}_slice"]) == (start, end)
        got = orc.prematch_utterance(start, end, pool, f0, spec, harm)
        ref_n = golden_pm[f"pm_u{u}_nearest_nbrs"]
        lo = np.full(ln, start); hi = np.full(ln, end)
        _, val33 = orc.knn(pool[start:end], pool, 33, lo, hi)
        assert not ((ref_n >= start) & (ref_n < end)).any()                   # own frames never selected
        m = positions_untied(val33, 32)
        assert m.mean() > 0.5
        assert np.array_equal(got["nearest_nbrs"][m], ref_n[m])
        rows = set_rows(val33, 32)
        assert np.array_equal(np.sort(got["nearest_nbrs"][rows], 1), np.sort(ref_n[rows], 1))
        # later stages from the reference's own top-32 (tied slots may legitimately differ above)
        got = orc.prematch_utterance(start, end, pool, f0, spec, harm, nbrs=ref_n)
        assert np.array_equal(got["nearest_nbrs_f0_priority"], golden_pm[f"pm_u{u}_nearest_nbrs_f0_priority"])
        ref_a = golden_pm[f"pm_u{u}_amp_ratio"]
        assert np.abs(got["amp_ratio"] - ref_a).max() <= 1e-5 * np.abs(ref_a).max()
        ref_w = golden_pm[f"pm_u{u}_harmonics_best_weight_para"]
        assert np.abs(got["harmonics_best_weight_para"] - ref_w).max() < 5e-2
        start = end
