"""Edge branches of the matcher path, from THE REFERENCE ITSELF (read-only import from /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_edges.py  ->  reference_outputs_edges.npz

 * greedy re-selection at the END of the pool (`prev + 1` clamped to Np-1, lib_ongaku_test.py:294-295),
   at other concat weights (post_opt_0.1, post_opt_extra = 0.3) and on a two-frame utterance;
 * f0 re-rank with unvoiced (0 Hz) query and pool frames (log2(0 + 1e-5) keys);
 * the f0 shift with unvoiced frames on both sides (ddsp_prematch_dataset.py:1224-1233);
 * harmonic bank at high f0: most harmonics above Nyquist are removed (:146-156), and an all-unvoiced track;
 * match_at_inference_time with ckpt_type="wavlm_only" (arity 3, topk-mean only).
Inputs are regenerated from seeds by knn_svc_b200/synth.py."""
import contextlib
import io
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True

import lib_ongaku_test as ref_lib          # noqa: E402  (reference)
import ddsp_prematch_dataset as ref_pm     # noqa: E402  (reference)
from knn_svc_b200 import synth             # noqa: E402

torch.manual_seed(0)
torch.set_num_threads(8)


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def main():
    out = {}
    # ---- K5 at the end of the pool: every candidate sits in the last rows, prev+1 clamps
    T, Np = 60, 300
    q = synth.ar1_frames(T, seed=111, reset_every=25)
    p = synth.ar1_frames(Np, seed=112)
    rs = np.random.RandomState(5)
    idx = np.sort(rs.randint(Np - 6, Np, size=(T, 4)), axis=1).astype(np.int64)
    idx[::7] = Np - 1                                              # rows whose four candidates are all the last row
    out["k5e_idx"] = idx
    f0q, f0p = synth.f0_track(T, seed=113, unvoiced=0.3), synth.f0_track(Np, seed=114, unvoiced=0.3)
    for w, tag in ((0.2, "w0p2"), (0.1, "w0p1"), (0.3, "w0p3")):
        for dt, dtag in ((torch.float32, "f32"), (torch.float64, "f64")):
            r = quiet(ref_lib.knn_with_concat_cost, torch.from_numpy(idx).clone(), torch.from_numpy(q).to(dt),
                      torch.from_numpy(p).to(dt), concat_weight=w)
            out[f"k5e_nof0_{tag}_{dtag}"] = r.numpy()
            r = quiet(ref_lib.knn_with_concat_cost, torch.from_numpy(idx).clone(), torch.from_numpy(q).to(dt),
                      torch.from_numpy(p).to(dt), torch.from_numpy(f0q), torch.from_numpy(f0p), concat_weight=w)
            out[f"k5e_f0_{tag}_{dtag}"] = r.numpy()
    r = quiet(ref_lib.knn_with_concat_cost, torch.from_numpy(idx[:2]).clone(), torch.from_numpy(q[:2]).double(),
              torch.from_numpy(p).double(), concat_weight=0.2)
    out["k5e_two_frames"] = r.numpy()

    # ---- K4 with unvoiced frames on both sides
    nb = np.stack([rs.permutation(Np)[:32] for _ in range(T)]).astype(np.int64)
    out["k4e_nbrs"] = nb
    out["k4e_prio"] = ref_pm.sort_by_f0_compatibility(torch.from_numpy(f0q), torch.from_numpy(f0p), torch.from_numpy(nb)).numpy()

    # ---- a6 f0 shift exactly as :1224-1233 (voiced medians, unvoiced frames stay 0)
    src_f0, tgt_f0 = torch.from_numpy(f0q).clone(), torch.from_numpy(f0p)
    voiced_src = src_f0[src_f0 != 0]
    voiced_tgt = tgt_f0[tgt_f0 != 0]
    shifted = src_f0.clone()
    shifted[src_f0 != 0] = torch.exp(torch.log(voiced_src) + torch.median(torch.log(voiced_tgt)) - torch.median(torch.log(voiced_src)))
    out["a6e_shifted"] = shifted.numpy()

    # ---- K7 at high f0 (harmonics above Nyquist removed) and on an all-unvoiced track
    T7 = 16
    f0hi = np.linspace(700.0, 1000.0, T7).astype(np.float32)
    f0hi[5:8] = 0.0
    amp = synth.harmonics_pool(T7, seed=115)
    out["k7e_hi"] = ref_pm.get_bulk_dsp_choral(torch.from_numpy(f0hi)[None, :, None], torch.from_numpy(amp)[None]).numpy()
    out["k7e_zero"] = ref_pm.get_bulk_dsp_choral(torch.zeros(1, T7, 1), torch.from_numpy(amp)[None]).numpy()

    # ---- a11 with ckpt_type="wavlm_only": arity 3, plain top-k mean, no harmonics
    Tq, Npp = 70, 250
    qf, pf = synth.ar1_frames(Tq, seed=116, reset_every=30), synth.ar1_frames(Npp, seed=117)
    f0a, f0b = synth.f0_track(Tq, seed=118), synth.f0_track(Npp, seed=119)
    hp = synth.harmonics_pool(Npp, seed=120)

    def fake_pool(wav, *a, **k):
        if "src" in str(wav):
            feats, f0, n, harm = torch.from_numpy(qf).double(), torch.from_numpy(f0a), Tq, torch.zeros(Tq, 49)
        else:
            feats, f0, n, harm = torch.from_numpy(pf).double(), torch.from_numpy(f0b), Npp, torch.from_numpy(hp)
        key = str(wav)
        return ({key: feats}, {key: feats}, {key: torch.zeros(n, 320)}, {key: torch.ones(n, 201)}, {key: f0}, {key: harm})

    ref_pm.get_complete_spk_pool = fake_pool
    res = quiet(ref_pm.match_at_inference_time, Path("/x/src.wav"), Path("/x/ref.wav"), None, None, None, device="cpu",
                prioritize_f0=True, ckpt_type="wavlm_only", src_dataset_path="/x", tgt_dataset_path="/x",
                post_opt="no_post_opt")
    out["a11e_arity"] = np.int64(len(res))
    out["a11e_feats"] = res[0]["/x/src.wav"].detach().numpy()[:, ::8]
    out["a11e_f0"] = res[-1]["/x/src.wav"].numpy()
    np.savez_compressed(HERE / "reference_outputs_edges.npz", **out)
    print({k: (np.asarray(v).shape, str(np.asarray(v).dtype)) for k, v in out.items()})


if __name__ == "__main__":
    main()
