"""The oracle against the reference's own outputs at BASELINE cfg 1 / cfg 2's real shape
(T = Np = 3001 frames, the real f0 tracks; tests/golden/make_golden_cfg12.py).  CPU only."""
from pathlib import Path

import numpy as np
import pytest

from knn_svc_b200 import synth
from oracle import matcher_oracle as orc
from tests.util import positions_untied, set_rows

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def g():
    return dict(np.load(ROOT / "tests" / "golden" / "reference_outputs_cfg12.npz"))


@pytest.fixture(scope="module")
def inputs(g):
    T = len(g["f0_src"])
    return synth.ar1_frames(T, seed=301, reset_every=200), synth.ar1_frames(T, seed=302), synth.harmonics_pool(T, seed=303)


def test_oracle_search_and_reselection_at_full_shape(g, inputs):
    qf, pf, hp = inputs
    o_idx, o_val = orc.knn(qf, pf, 33)
    ref_idx, ref_val = g["nbrs33"].astype(np.int64), g["vals33"]
    assert np.abs(o_val - ref_val).max() < 2e-6
    m = positions_untied(ref_val, 32)
    assert m.mean() > 0.5 and np.array_equal(o_idx[:, :32][m], ref_idx[:, :32][m])
    rows = set_rows(ref_val, 32)
    assert np.array_equal(np.sort(o_idx[rows, :32], 1), np.sort(ref_idx[rows, :32], 1))
    shifted = orc.shift_f0(g["f0_src"], g["f0_tgt"])
    assert np.abs(shifted - g["no_post_opt_f0"]).max() <= 1e-6 * g["no_post_opt_f0"].max()
    prio = orc.sort_by_f0_compatibility(shifted, g["f0_tgt"], ref_idx[:, :32])
    assert np.array_equal(prio, g["prio32"])
    sel = orc.knn_with_concat_cost(ref_idx[:, :4], qf, pf, concat_weight=0.2)
    assert np.array_equal(sel, g["k5_nof0"])
    sel_f0 = orc.knn_with_concat_cost(g["prio32"][:, :4].astype(np.int64), qf, pf, shifted, g["f0_tgt"], concat_weight=0.2)
    assert np.array_equal(sel_f0, g["k5_f0"])


def test_oracle_no_post_opt_outputs_at_full_shape(g, inputs):
    qf, pf, hp = inputs
    ref_idx = g["nbrs33"].astype(np.int64)
    feats = orc.gather_mix(pf, ref_idx[:, :4], None)
    ref = g["no_post_opt_feats_sub"]
    assert np.abs(feats[:, ::16] - ref).max() <= 1e-6 * np.abs(ref).max()
    harm = orc.gather_mix(hp, g["prio32"][:, :4].astype(np.int64), None)
    assert np.abs(harm - g["no_post_opt_harm"]).max() <= 1e-6 * np.abs(g["no_post_opt_harm"]).max()
