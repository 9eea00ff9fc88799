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
