import numpy as np

GAP = 1e-5  # north-star tie criterion: indices must match where distances differ by more than this


def positions_untied(vals, k):
    """[T,k] mask: slot j is separated from both sorted neighbours by > GAP
    (vals: [T,k+1] sorted ascending distances from the oracle)."""
    d = np.diff(vals[:, :k + 1], axis=1) > GAP
    before = np.concatenate([np.ones((len(vals), 1), bool), d[:, :-1]], axis=1)
    return d & before


def set_rows(vals, k):
    """rows whose k-th/(k+1)-th gap exceeds GAP: the top-k SET is determined"""
    return (vals[:, k] - vals[:, k - 1]) > GAP


def check_knn_against_oracle(idx, dist, o_idx, o_val, k, min_cover=0.3):
    """idx/dist [T,k] from the device; o_idx/o_val [T,k+1] from the oracle."""
    idx, dist = np.asarray(idx), np.asarray(dist)
    assert idx.shape == (len(o_idx), k)
    assert np.abs(dist - o_val[:, :k]).max() < 2e-6, np.abs(dist - o_val[:, :k]).max()
    m = positions_untied(o_val, k)
    assert m.mean() >= min_cover, m.mean()
    assert np.array_equal(idx[m], o_idx[:, :k][m]), "top-k index mismatch on an untied slot"
    rows = set_rows(o_val, k)
    assert np.array_equal(np.sort(idx[rows], 1), np.sort(o_idx[rows, :k], 1)), "top-k set mismatch"
    return m.mean(), rows.mean()


def prematch_inputs(golden_pm):
    """Seeded synthetic speaker (3 utterances) of tests/golden/make_golden_prematch.py;
    the f0 track is real data and comes from the fixture."""
    from knn_svc_b200 import synth
    lens = [60, 90, 50]
    feats = synth.ar1_frames(sum(lens), seed=61, reset_every=70)
    spec = np.abs(synth.randn_frames(sum(lens), 200, seed=62)).astype(np.float32) + 0.05
    harm = synth.harmonics_pool(sum(lens), seed=63)
    return lens, feats, spec, harm, golden_pm["pm_f0"]


def pool_builder_inputs(golden_pm):
    """audio crop (float32 in [-1,1)), layer features [25,T,64], f0 of the pool-builder fixture"""
    from knn_svc_b200 import synth
    x = golden_pm["pb_pcm"].astype(np.float32) / 32768.0
    T = golden_pm["pb_spec"].shape[0]
    layer_feats = synth.randn_frames(25 * T, 64, seed=71).reshape(25, T, 64)
    return x, layer_feats, golden_pm["pb_f0"][:T]


class FakeWavLM:
    """Deterministic stand-in for the WavLM encoder (the checkpoint is not available): conv-stack
    framing (400-sample window, hop 320) followed by fixed per-layer projections.  Used by the
    fixture script with the REFERENCE's KNeighborsVC and by the tests with ours."""

    class cfg:
        encoder_layers = 24

    def __init__(self, dim=16, seed=5):
        import torch
        g = torch.Generator().manual_seed(seed)
        self.proj = torch.randn((25, 400, dim), generator=g) * 0.05
        self.calls = []

    def eval(self):
        return self

    def extract_features(self, wav, output_layer=None, ret_layer_results=False):
        import torch
        self.calls.append((tuple(wav.shape), output_layer, ret_layer_results))
        frames = wav[0].unfold(0, 400, 320)                         # [T, 400]
        layers = [torch.tanh(frames @ self.proj[l]) for l in range(25)]      # 25 x [T, dim]
        if ret_layer_results:
            return ((layers[-1][None], [(l[:, None, :], None) for l in layers]), None)
        return (layers[output_layer][None], None)


def fake_waveform(n=16000 * 2, seed=3):
    import torch
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(n) / 16000.0
    return (0.3 * torch.sin(2 * np.pi * 220.0 * t) + 0.02 * torch.randn(n, generator=g)).float()
