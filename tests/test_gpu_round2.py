"""Round-2 parity additions (VERDICT r1 "next round" 1b-1d, weak 5), through the C ABI:

  * the tensor-core accumulator values themselves (read back from the filter's candidate log) are
    inside the error window the search relies on, on adversarial rows, dim 1024 and 4096, fp16 and
    bf16 operands;
  * BASELINE cfg 1 / cfg 2 at their real shape (T = Np = 3001, real f0 tracks) against the
    reference's own outputs (tests/golden/make_golden_cfg12.py);
  * KNeighborsVC.match(post_opt=...) checked for VALUES against the oracle's composition;
  * gather_mix over a sharded row table == gather_mix over the contiguous pool (bit for bit).
"""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from knn_svc_b200 import synth
from oracle import matcher_oracle as orc
from tests.util import GAP, check_knn_against_oracle, positions_untied, set_rows

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = Path(__file__).resolve().parent.parent


def dev(x, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(x)).to(DEV)
    return t if dtype is None else t.to(dtype)


@pytest.fixture(scope="module")
def ops():
    from knn_svc_b200 import ops as _ops
    assert torch.cuda.is_available()
    return _ops


def _set_opt(name, value):
    from knn_svc_b200 import _lib
    _lib.check(_lib.load().knnsvc_set_option(name.encode(), int(value)), "set_option")


def _adversarial_rows(n, d, seed):
    """rows that stress the accumulator: all-positive mean-shifted (WavLM-like: every product has
    the same sign, partial sums grow monotonically to ~1), heavy-tailed, near-duplicates of each
    other (s ~ 1), sign-flipped copies (s ~ -1), a few huge components, tiny components"""
    rs = np.random.RandomState(seed)
    x = np.abs(rs.standard_normal((n, d))) + 3.0                       # all positive, common mean
    x[n // 4: n // 2] = rs.standard_t(2.0, size=(n // 2 - n // 4, d))  # heavy tails
    x[n // 2: n // 2 + 8] = x[:8] * (1 + 1e-3 * rs.standard_normal((8, d)))       # near duplicates
    x[n // 2 + 8: n // 2 + 16] = -x[:8]                                # anti-parallel
    x[n // 2 + 16: n // 2 + 24, :4] *= 300.0                           # energy in a few components
    x[n // 2 + 24: n // 2 + 32] = 1e-3 * rs.standard_normal((8, d)) + 1.0          # nearly constant rows
    return x.astype(np.float32)


@pytest.mark.parametrize("fmt", ["fp16", "bf16"])
@pytest.mark.parametrize("dim", [1024, 4096])
def test_filter_accumulator_error_is_inside_the_window(ops, dim, fmt):
    """|s~ - s| <= eps with s~ the value tcgen05.mma left in TMEM (as logged by the filter's
    epilogue), s the exact fp64 cosine similarity and eps the window the search uses — and the
    accumulator's own share |s~ - (u_q . u_p)| <= acc_eps(dim), u the fp16/bf16 operands.  A pool of
    k = 32 rows makes the filter log EVERY column (its threshold only exists once k values were seen),
    so the whole accumulator tile is read back, chunk by chunk."""
    _set_opt("bf16_operands", fmt == "bf16")
    try:
        q = _adversarial_rows(256, dim, seed=1)
        p = np.concatenate([_adversarial_rows(192, dim, seed=2), q[:64] * 0.5])      # incl. exact parallels
        qp, pp_all = ops.prepare_rows(dev(q)), ops.prepare_rows(dev(p))
        scale = 1024.0
        uq = (qp.half.view(torch.bfloat16) if fmt == "bf16" else qp.half).double().cpu().numpy()[:, :dim] / scale
        unit_q = q.astype(np.float64) / np.linalg.norm(q.astype(np.float64), axis=1, keepdims=True)
        acc_eps = 1.4e-4 * max(1.0, dim / 1024.0)
        worst_total, worst_acc, n_vals = 0.0, 0.0, 0
        for a in range(0, len(p), 32):
            chunk = p[a:a + 32]
            pp = ops.prepare_rows(dev(chunk))
            val, col, cnt, lay, _ = ops.knn_candidate_log(qp, pp, 32)
            assert lay["n_seg"] == 1 and int(cnt.min()) == 32 and int(cnt.max()) == 32, "every column must be logged"
            val, col = val[:, :32].double().cpu().numpy(), col[:, :32].cpu().numpy().astype(np.int64)
            up = (pp.half.view(torch.bfloat16) if fmt == "bf16" else pp.half).double().cpu().numpy()[:, :dim] / scale
            unit_p = chunk.astype(np.float64) / np.linalg.norm(chunk.astype(np.float64), axis=1, keepdims=True)
            s_exact = np.take_along_axis(unit_q @ unit_p.T, col, axis=1)
            s_oper = np.take_along_axis(uq @ up.T, col, axis=1)
            eps = float(qp.err.item()) + float(pp.err.item()) + float(qp.err.item()) * float(pp.err.item()) + acc_eps
            worst_total = max(worst_total, float(np.abs(val - s_exact).max() / eps))
            worst_acc = max(worst_acc, float(np.abs(val - s_oper).max()))
            n_vals += val.size
        print(f"accumulator check dim={dim} {fmt}: {n_vals} values, max |s~-s|/eps = {worst_total:.3f}, "
              f"max |s~ - u_q.u_p| = {worst_acc:.2e} (bound {acc_eps:.1e})")
        out = ROOT / "gpurun_out"
        out.mkdir(exist_ok=True)
        with open(out / "r2_accumulator_check.jsonl", "a") as f:
            f.write(json.dumps({"dim": dim, "operands": fmt, "values": n_vals, "max_err_over_eps": worst_total,
                                "max_accumulator_err": worst_acc, "acc_eps_bound": acc_eps}) + "\n")
        assert worst_total <= 1.0
        assert worst_acc <= acc_eps
    finally:
        _set_opt("bf16_operands", 0)


def test_filter_log_holds_every_true_neighbour_with_its_accumulator_value(ops):
    """on a WavLM-like set at a realistic size: every true top-32 member is in the log, and each
    logged s~ is within the measured eps of the exact similarity"""
    q, p = synth.ar1_frames(300, seed=41, reset_every=100), synth.ar1_frames(30000, seed=42)
    qp, pp = ops.prepare_rows(dev(q)), ops.prepare_rows(dev(p))
    val, col, cnt, lay, res = ops.knn_candidate_log(qp, pp, 32)
    assert int(res[2][0]) == 0                          # no row needed the exact fallback
    n_seg, cap = lay["n_seg"], lay["cap"]
    val, col, cnt = val.cpu().numpy(), col.cpu().numpy(), cnt.cpu().numpy()
    o_idx, o_val = orc.knn(q, p, 33)
    unit_q = q.astype(np.float64) / np.linalg.norm(q.astype(np.float64), axis=1, keepdims=True)
    unit_p = p.astype(np.float64) / np.linalg.norm(p.astype(np.float64), axis=1, keepdims=True)
    eps = float(qp.err.item()) + float(pp.err.item()) + float(qp.err.item()) * float(pp.err.item()) + 1.4e-4
    worst = 0.0
    for t in range(len(q)):
        logged = {}
        for s in range(n_seg):
            n = min(int(cnt[t * n_seg + s]), cap)
            logged.update(zip(col[t * n_seg + s, :n].tolist(), val[t * n_seg + s, :n].tolist()))
        assert set(o_idx[t, :32].tolist()) <= set(logged), f"row {t}: a true neighbour is missing from the log"
        cols = np.fromiter(logged.keys(), dtype=np.int64)
        s_exact = unit_p[cols] @ unit_q[t]
        worst = max(worst, float(np.abs(np.fromiter(logged.values(), dtype=np.float64) - s_exact).max()))
    print(f"logged s~ vs exact on 300x30000 AR(1): max |s~ - s| = {worst:.2e}, eps = {eps:.2e}")
    assert worst <= eps


# ----------------------------------------------------------------------------- cfg 1 / cfg 2, real shape
@pytest.fixture(scope="module")
def golden12():
    return dict(np.load(ROOT / "tests" / "golden" / "reference_outputs_cfg12.npz"))


def _cfg12_inputs(g):
    T = len(g["f0_src"])
    qf = synth.ar1_frames(T, seed=301, reset_every=200)
    pf = synth.ar1_frames(T, seed=302)
    hp = synth.harmonics_pool(T, seed=303)
    return qf, pf, g["f0_src"], g["f0_tgt"], hp


def test_cfg12_search_and_reselection_match_the_reference(ops, golden12):
    """T = Np = 3001: top-32 indices, f0 re-rank and BOTH greedy re-selections over all 3001 frames
    (the recurrence is compared up to the first legitimately tied row, of which there is none here)"""
    from knn_svc_b200 import ddsp_prematch_dataset as pm
    qf, pf, f0q, f0p, hp = _cfg12_inputs(golden12)
    qp, pp = ops.prepare_rows(dev(qf)), ops.prepare_rows(dev(pf))
    dist, idx = ops.knn_search(qp, pp, 32)
    o_idx, o_val = golden12["nbrs33"].astype(np.int64), golden12["vals33"]
    check_knn_against_oracle(idx.cpu().numpy(), dist.cpu().numpy(), o_idx, o_val, 32)
    # downstream stages are fed the REFERENCE's indices so that each stage is pinned on its own
    ref32 = dev(o_idx[:, :32])
    shifted = pm.shift_query_f0(torch.from_numpy(f0q), torch.from_numpy(f0p))
    prio = pm.sort_by_f0_compatibility(shifted, dev(f0p), ref32).cpu().numpy()
    assert np.array_equal(prio, golden12["prio32"])
    sel = pm.knn_with_concat_cost(ref32[:, :4].contiguous(), qp.rows, pp.rows, concat_weight=0.2).cpu().numpy()
    assert np.array_equal(sel, golden12["k5_nof0"]), int((sel != golden12["k5_nof0"]).any(1).argmax())
    sel_f0 = pm.knn_with_concat_cost(dev(golden12["prio32"][:, :4].astype(np.int64)), qp.rows, pp.rows, shifted.to(DEV),
                                     dev(f0p), concat_weight=0.2).cpu().numpy()
    assert np.array_equal(sel_f0, golden12["k5_f0"]), int((sel_f0 != golden12["k5_f0"]).any(1).argmax())


@pytest.mark.parametrize("post_opt", ["no_post_opt", "post_opt_0.2"])
def test_cfg12_pipeline_matches_the_reference(ops, golden12, post_opt):
    """match_at_inference_time at cfg 1 / cfg 2 shape against the reference's run of the same call"""
    from knn_svc_b200 import ddsp_prematch_dataset as pm
    qf, pf, f0q, f0p, hp = _cfg12_inputs(golden12)

    def fake_pool(wav, *a, **k):
        if "src" in str(wav):
            feats, f0, n, harm = torch.from_numpy(qf).double(), torch.from_numpy(f0q), len(qf), torch.zeros(len(qf), 49)
        else:
            feats, f0, n, harm = torch.from_numpy(pf).double(), torch.from_numpy(f0p), len(pf), torch.from_numpy(hp)
        key = str(wav)
        return ({key: feats}, {key: feats}, {key: torch.zeros(n, 320)}, {key: torch.ones(n, 201)}, {key: f0}, {key: harm})

    old = pm.get_complete_spk_pool
    pm.get_complete_spk_pool = fake_pool
    try:
        feats, harm, audio, sf0 = pm.match_at_inference_time(
            Path("/x/src.wav"), Path("/x/ref.wav"), None, None, None, device=DEV, prioritize_f0=True,
            ckpt_type="mix", src_dataset_path="/x", tgt_dataset_path="/x", post_opt=post_opt)
    finally:
        pm.get_complete_spk_pool = old
    tag = post_opt.replace(".", "p")
    fe = feats["/x/src.wav"].cpu().numpy()
    ref = golden12[f"{tag}_feats_sub"]
    ha, ref_h = harm["/x/src.wav"].cpu().numpy(), golden12[f"{tag}_harm"]
    # rows whose 4th / 5th reference distances differ by more than 1e-5 carry the north-star gate;
    # the fp64 re-score (fp64 norms) in fact reproduces the reference's order on ALL rows here
    # (its closest 4th/5th pair is 2.2e-8 apart), which is recorded.
    vals = golden12["vals33"]
    decided = (vals[:, 4] - vals[:, 3]) > GAP
    rel_all = np.abs(fe[:, ::16] - ref).max() / np.abs(ref).max()
    rel = np.abs(fe[decided][:, ::16] - ref[decided]).max() / np.abs(ref).max()
    rel_sum = np.abs(fe.astype(np.float64).sum(1) - golden12[f"{tag}_feats_rowsum"])[decided].max() / np.abs(golden12[f"{tag}_feats_rowsum"]).max()
    relh = np.abs(ha - ref_h).max() / np.abs(ref_h).max()
    relf = np.abs(sf0["/x/src.wav"].numpy() - golden12[f"{tag}_f0"]).max() / golden12[f"{tag}_f0"].max()
    print(f"cfg1/2 pipeline {post_opt} (3001 x 3001): feats rel {rel:.2e}, row sums rel {rel_sum:.2e}, "
          f"harmonics rel {relh:.2e}, f0 rel {relf:.2e}")
    with open(ROOT / "gpurun_out" / "r2_cfg12_deviation.jsonl", "a") as f:
        f.write(json.dumps({"post_opt": post_opt, "feats_rel": float(rel), "feats_rel_all_rows_incl_ties": float(rel_all),
                            "rows_with_4th_5th_gap_below_1e-5": int((~decided).sum()), "rowsum_rel": float(rel_sum),
                            "harmonics_rel": float(relh), "f0_rel": float(relf)}) + "\n")
    assert relf <= 1e-6
    if post_opt == "no_post_opt":
        assert rel <= 1e-4 and relh <= 1e-4 and rel_sum <= 1e-4      # north-star tolerance
    else:
        assert rel < 5e-3 and relh < 5e-2                            # passes through the Adam fit


# ----------------------------------------------------------------------------- match(post_opt) values
def test_matcher_match_post_opt_values(ops):
    """KNeighborsVC.match(post_opt="post_opt_0.2") == oracle composition: kNN -> greedy re-selection
    -> fitted mixing weights -> mix.  Indices exact; the features inherit the Adam fit's conditioning
    (SURVEY D13), so they are gated through the achieved smoothness loss and a loose bound."""
    from knn_svc_b200.ddsp_matcher import KNeighborsVC
    m = KNeighborsVC(None, None, None, device=DEV)
    q, p = synth.ar1_frames(160, seed=171, reset_every=60), synth.ar1_frames(900, seed=172)
    out = m.match(torch.from_numpy(q), torch.from_numpy(p), topk=4, without_vocode=True, post_opt="post_opt_0.2")
    o_idx, o_val = orc.knn(q, p, 5)
    # the search itself is pinned elsewhere; rows whose 4th/5th distances tie within 1e-5 may legitimately
    # pick another neighbour and the greedy recurrence would carry that on, so the oracle's composition
    # starts from the device's own top-4 (checked against the oracle on the decided rows)
    _, top4 = ops.knn_search(ops.prepare_rows(dev(q)), ops.prepare_rows(dev(p)), 4)
    top4 = top4.cpu().numpy()
    rows = set_rows(o_val, 4)
    assert rows.mean() > 0.8 and np.array_equal(np.sort(top4[rows], 1), np.sort(o_idx[rows, :4], 1))
    sel = orc.knn_with_concat_cost(top4, q, p, concat_weight=0.2)
    w = orc.compute_wavlm_weight(sel, p)
    want = orc.gather_mix(p, sel, w)
    got = out.cpu().numpy()
    rel = np.abs(got - want).max() / np.abs(want).max()
    rows = orc._neighbour_rows(sel, np.asarray(p, np.float64))
    # recover our weights' loss from the features is not possible; compare the loss of the oracle's fit with
    # the loss of OUR fit on the same indices
    from knn_svc_b200 import ddsp_prematch_dataset as pm
    w_ours = pm.compute_wavlm_weight(dev(sel), dev(p)).cpu().numpy().astype(np.float64)
    l_ref, l_got = orc.smoothness_loss(w.astype(np.float64), rows, 0.1), orc.smoothness_loss(w_ours, rows, 0.1)
    print(f"match(post_opt_0.2): feats rel {rel:.2e}; loss ref {l_ref:.9e} ours {l_got:.9e}")
    assert abs(l_ref - l_got) <= 1e-6 * abs(l_ref) + 1e-9
    assert rel < 5e-3
    # and the no-fit / no-reselect corner: post_opt string that parses to -1 but is not "no_post_opt"
    out3 = m.match(torch.from_numpy(q), torch.from_numpy(p), topk=4, without_vocode=True, post_opt="no_post_opt")
    assert np.abs(out3.cpu().numpy() - orc.gather_mix(p, top4, None)).max() <= 1e-4 * np.abs(want).max()


# ----------------------------------------------------------------------------- dense decision route
@pytest.mark.parametrize("k", [4, 32])
def test_refine_route_on_a_near_duplicate_pool(ops, k):
    """a pool of near-duplicates (every row has ~500 copies within 3e-4 cosine) puts > 400 candidates per
    row inside the fp16 window: the decision stage takes its refine route (fp32 re-score, block-major), which
    must return exactly what the exact CUDA-core kernel, the direct route and the oracle return — masked
    ranges included"""
    rs = np.random.RandomState(5)
    base = synth.ar1_frames(48, seed=91)
    pool = (np.tile(base, (500, 1)) + 0.05 * rs.standard_normal((24000, 1024))).astype(np.float32)
    q = (np.repeat(synth.ar1_frames(48, seed=91), 6, axis=0)[:270] + 0.05 * rs.standard_normal((270, 1024))).astype(np.float32)
    qp, pp = ops.prepare_rows(dev(q)), ops.prepare_rows(dev(pool))
    lo = dev(np.full(270, 1000, np.int64)); hi = dev(np.full(270, 9000, np.int64))
    de, ie = ops.knn_exact(qp, pp, k)
    d, i, st = ops.knn_search(qp, pp, k, return_stats=True)
    dm, im = ops.knn_search(qp, pp, k, mask_lo=lo, mask_hi=hi)
    assert int(st[2]) > 400 * 270, "fixture: the refine route needs > 400 candidates per row"
    assert int(st[0]) == 0 and int(st[7]) < int(st[2]) // 4, "the fp32 stage must thin the candidates out"
    assert torch.equal(i, ie) and torch.equal(d, de)
    # the direct route (threshold raised out of reach) must return the same bits
    _set_opt("refine_min_candidates", 1 << 20)
    try:
        d2, i2, st2 = ops.knn_search(qp, pp, k, return_stats=True)
    finally:
        _set_opt("refine_min_candidates", 0)
    assert int(st2[7]) == int(st2[2]) and torch.equal(i2, i) and torch.equal(d2, d)
    outs = {1: (d, i, dm, im)}
    o_idx, o_val = orc.knn(q, pool, k + 1)
    check_knn_against_oracle(outs[1][1].cpu().numpy(), outs[1][0].cpu().numpy(), o_idx, o_val, k, min_cover=0.0)
    om_idx, om_val = orc.knn(q, pool, k + 1, np.full(270, 1000), np.full(270, 9000))
    check_knn_against_oracle(outs[1][3].cpu().numpy(), outs[1][2].cpu().numpy(), om_idx, om_val, k, min_cover=0.0)


@pytest.mark.parametrize("k", [4, 32])
def test_direct_route_with_dozens_of_candidates_per_row(ops, k):
    """rows with a dozen to a few dozen candidates inside the fp16 window (a pool with 10 near-duplicates of every
    row: the cfg-5 regime) take the direct route — one warp per row up to 32 candidates, one CTA per row above:
    same bits as the exact kernel, masked ranges included, a row count that is not a multiple of anything"""
    rs = np.random.RandomState(8)
    base = synth.ar1_frames(600, seed=93)
    pool = (np.tile(base, (10, 1)) + 0.05 * rs.standard_normal((6000, 1024))).astype(np.float32)
    q = (base[:523] + 0.05 * rs.standard_normal((523, 1024))).astype(np.float32)
    qp, pp = ops.prepare_rows(dev(q)), ops.prepare_rows(dev(pool))
    lo = dev(np.full(523, 1200, np.int64)); hi = dev(np.full(523, 3100, np.int64))
    de, ie = ops.knn_exact(qp, pp, k)
    d, i, st = ops.knn_search(qp, pp, k, return_stats=True)
    dm, im = ops.knn_search(qp, pp, k, mask_lo=lo, mask_hi=hi)
    per_row = int(st[2]) / 523
    print(f"k={k}: {per_row:.1f} candidates per row, {int(st[7])} scored in fp64, flagged {int(st[0])}")
    # (k = 32: ~50 per row, one CTA per row; k = 4: ~10 per row, one warp per row)
    assert (33 if k == 32 else 4) < per_row < 256 and int(st[0]) == 0 and int(st[7]) == int(st[2]), "fixture: direct route"
    assert torch.equal(i, ie) and torch.equal(d, de)
    om_idx, om_val = orc.knn(q, pool, k + 1, np.full(523, 1200), np.full(523, 3100))
    check_knn_against_oracle(im.cpu().numpy(), dm.cpu().numpy(), om_idx, om_val, k, min_cover=0.0)


def test_decision_route_is_chosen_per_row(ops):
    """one search whose first half of the query rows sits in a pool of near-duplicates (> 400 candidates inside
    the fp16 window: refine route) and whose second half is i.i.d. (a handful of candidates: direct route,
    decided by one warp per row): every row takes the route it would take alone — the candidate counts of the
    mixed search are the sums of the two halves searched separately — and the result is the exact kernel's"""
    rs = np.random.RandomState(6)
    base = synth.ar1_frames(48, seed=92)
    q_iid = rs.standard_normal((135, 1024)).astype(np.float32)
    # the i.i.d. rows find 8 loose copies of themselves (cosine 0.89 +- 0.014: far wider than the window)
    pool = np.concatenate([np.tile(base, (500, 1)) + 0.05 * rs.standard_normal((24000, 1024)),
                           np.repeat(q_iid, 8, axis=0) + 0.5 * rs.standard_normal((1080, 1024)),
                           rs.standard_normal((6920, 1024))]).astype(np.float32)
    q_dense = (np.repeat(base, 3, axis=0)[:135] + 0.05 * rs.standard_normal((135, 1024))).astype(np.float32)
    # interleave in runs of 5 rows so that 32-row groups of the refine kernel hold rows of both routes
    order = np.argsort(np.concatenate([np.arange(135) // 5 * 2, np.arange(135) // 5 * 2 + 1]), kind="stable")
    q = np.concatenate([q_dense, q_iid])[order]
    pp = ops.prepare_rows(dev(pool))
    k = 4
    d, i, st = ops.knn_search(ops.prepare_rows(dev(q)), pp, k, return_stats=True)
    de, ie = ops.knn_exact(ops.prepare_rows(dev(q)), pp, k)
    assert torch.equal(i, ie) and torch.equal(d, de)
    _, _, st_a = ops.knn_search(ops.prepare_rows(dev(q_dense)), pp, k, return_stats=True)
    _, _, st_b = ops.knn_search(ops.prepare_rows(dev(q_iid)), pp, k, return_stats=True)
    st, st_a, st_b = (x.cpu().numpy().astype(np.int64) for x in (st, st_a, st_b))
    print("mixed", st[:3], st[7], "dense", st_a[:3], st_a[7], "iid", st_b[:3], st_b[7])
    assert st[0] == st_a[0] == st_b[0] == 0
    assert st_a[2] > 400 * 135 and st_a[7] < st_a[2] // 4, "fixture: the dense half must take the refine route"
    assert st_b[2] <= 32 * 135 and st_b[7] == st_b[2], "fixture: the i.i.d. half must take the direct route"
    # (the operand error bound eps is the max over the row set, so it can differ by a hair between the three
    # query sets; the window, and with it the candidate count, is compared with 2% slack)
    assert abs(int(st[2]) - int(st_a[2] + st_b[2])) <= 0.02 * st[2]
    assert abs(int(st[7]) - int(st_a[7] + st_b[7])) <= 0.02 * st[7] + 8


# ----------------------------------------------------------------------------- sharded row table (one GPU)
@pytest.mark.parametrize("dim", [1024, 49])
def test_gather_mix_sharded_equals_contiguous(ops, dim):
    rs = np.random.RandomState(3)
    pool = synth.randn_frames(5000, d=dim, seed=5)
    bounds = [0, 1, 1700, 1700, 3333, 5000]                    # a 1-row shard and an empty shard
    pool_t = dev(pool)
    parts = [pool_t[a:b].clone() if b > a else pool_t[:1].clone() for a, b in zip(bounds[:-1], bounds[1:])]
    table = ops.ShardedRows([t.data_ptr() for t in parts], bounds, dim, torch.device(DEV))
    idx = rs.randint(0, 5000, size=(777, 4)).astype(np.int64)
    idx[0] = [0, 1, 1699, 1700]; idx[1] = [3332, 3333, 4999, 0]
    w = rs.rand(777, 4).astype(np.float32)
    for weights in (None, dev(w)):
        a = ops.gather_mix(pool_t, dev(idx), weights)
        b = ops.gather_mix_sharded(table, dev(idx), weights)
        assert torch.equal(a, b)
    assert ops.gather_mix_sharded(table, dev(idx[:0]), None).shape == (0, dim)


def test_post_opt_stage_on_a_row_table_equals_contiguous(ops):
    """K5 (both kernels, with and without f0) and K6 addressed through a 3-block row table return the
    bits they return on the contiguous pool; the blocks are cut so that `previous selection + 1` and
    the fit's idx +- 1 rows cross block boundaries all the time"""
    from knn_svc_b200 import ddsp_prematch_dataset as pm
    rs = np.random.RandomState(11)
    pool = synth.ar1_frames(1200, seed=61)
    q = synth.ar1_frames(300, seed=62, reset_every=90)
    f0q, f0p = synth.f0_track(300, seed=63), synth.f0_track(1200, seed=64)
    bounds = [0, 401, 402, 1200]
    pool_t = dev(pool)
    parts = [pool_t[a:b].clone() for a, b in zip(bounds[:-1], bounds[1:])]
    table = ops.ShardedRows([t.data_ptr() for t in parts], bounds, 1024, torch.device(DEV))
    base = rs.randint(395, 408, size=(300, 1))                       # candidates hug the block boundaries
    idx = np.clip(base + rs.randint(-3, 4, size=(300, 4)), 0, 1199).astype(np.int64)
    idx[250:] = rs.randint(0, 1200, size=(50, 4))
    idx[-1] = 1199                                                   # clamp at the end of the pool
    offs = [0, 120, 121, 300]
    for staged, clustered in ((1, 1), (1, 0), (0, 0)):       # cluster kernel, one-CTA staged kernel, general kernel
        _set_opt("concat_staged", staged); _set_opt("concat_cluster", clustered)
        try:
            for f0 in (None, (dev(f0q), dev(f0p))):
                args = () if f0 is None else f0
                a = ops.concat_cost_reselect(dev(idx), dev(q), pool_t, *args, concat_weight=0.2, utt_offsets=offs)
                b = ops.concat_cost_reselect(dev(idx), dev(q), table, *args, concat_weight=0.2, utt_offsets=offs)
                assert torch.equal(a, b), (staged, clustered, f0 is None)
        finally:
            _set_opt("concat_staged", 1); _set_opt("concat_cluster", 1)
    wa, ia = ops.weight_fit(dev(idx), pool_t, 0.1, return_info=True, utt_offsets=offs)
    wb, ib = ops.weight_fit(dev(idx), table, 0.1, return_info=True, utt_offsets=offs)
    assert torch.equal(wa, wb) and torch.equal(ia, ib)


def test_merge_topk64_ranks_on_fp64(ops):
    """two fp64 distances that round to the SAME fp32 value: the fp64 merge keeps their order, and
    shards that hold exact duplicates resolve to the lower global index"""
    a, b = 0.25, 0.25 + 2.0 ** -40
    gd = torch.tensor([[[b, 0.9]], [[a, 0.8]]], dtype=torch.float64, device=DEV)      # [R=2, T=1, k=2]
    gi = torch.tensor([[[5, 6]], [[70, 80]]], dtype=torch.int64, device=DEV)
    d, d64, i = ops.merge_topk64(gd, gi)
    assert i.tolist() == [[70, 5]] and d64.tolist() == [[a, b]] and d[0, 0] == d[0, 1]
    gd = torch.tensor([[[a, 0.9]], [[a, 0.8]]], dtype=torch.float64, device=DEV)
    d, d64, i = ops.merge_topk64(gd, gi)
    assert i.tolist() == [[5, 70]]


# ----------------------------------------------------------------------------- host reads that bypass the copy engine
def test_validation_flag_and_small_downloads_do_not_use_the_copy_engine(ops):
    """prepare_rows' bad-row counter lives in pinned host memory (the kernel writes it only for a bad row; reading
    it is a stream synchronisation, not a device-to-host copy), and to_host_small stores a small result into pinned
    staging memory by a kernel: neither waits behind a large download queued on another stream.  Functionally: bad
    rows are still refused (zero-norm, NaN, inf), good rows pass, and the download equals .cpu() bit for bit."""
    x = synth.ar1_frames(300, seed=5)
    ops.prepare_rows(dev(x))                                        # passes
    for bad_val, rows in ((0.0, [7]), (np.nan, [0, 299]), (np.inf, [150])):
        y = x.copy()
        for r in rows:
            y[r] = bad_val if bad_val == 0.0 else y[r]
            if bad_val != 0.0:
                y[r, 3] = bad_val
        with pytest.raises(ValueError, match=f"{len(rows)} zero-norm or non-finite"):
            ops.prepare_rows(dev(y))
    ops.prepare_rows(dev(x))                                        # the flag is reset for the next call
    ops.prepare_rows(dev(y), check=False)                           # unchecked: no error, as before
    g = torch.Generator(device=DEV); g.manual_seed(3)
    for shape, dtype in (((1,), torch.float32), ((1000, 3), torch.float32), ((262145,), torch.float32),
                         ((4097,), torch.float64), ((33, 4), torch.int64), ((5,), torch.int32)):
        t = (torch.randn(shape, device=DEV, generator=g) * 1000).to(dtype)
        h = ops.to_host_small(t)
        assert h.device.type == "cpu" and h.dtype == dtype and h.shape == t.shape and torch.equal(h, t.cpu())
    # while 1 GB is being downloaded on another stream, the small read returns long before that copy ends
    big = torch.empty((256 << 20,), dtype=torch.float32, device=DEV)
    big_host = torch.empty((256 << 20,), dtype=torch.float32).pin_memory()
    side = torch.cuda.Stream()
    small = torch.arange(1024, device=DEV, dtype=torch.float32)
    torch.cuda.synchronize()
    import time
    done = torch.cuda.Event()
    with torch.cuda.stream(side):
        big_host.copy_(big, non_blocking=True)
        done.record()
    t0 = time.perf_counter()
    h = ops.to_host_small(small)
    t_small = time.perf_counter() - t0
    still_copying = not done.query()
    done.synchronize()
    t_big = time.perf_counter() - t0
    print(f"small read {1e3 * t_small:.2f} ms while a {1e3 * t_big:.1f} ms download was in flight (still copying: {still_copying})")
    assert torch.equal(h, small.cpu())
    assert still_copying and t_small < 0.5 * t_big, "the small read waited for the download"
