"""Generate golden fixtures by running THE REFERENCE ITSELF (read-only import
from /root/reference) on seeded synthetic inputs.  Run in the build container:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference cannot travel to the GPU box, so the outputs are committed as
small .npz files next to this script.  Inputs are NOT stored: they are
regenerated from seeds by `knn_svc_b200/synth.py` (the two real f0 tracks of
`sample_content/` are stored, cropped, because they are data, 12 KB each).
Large outputs are stored as a column subsample plus per-row sums.
"""
import contextlib
import io
import os
import re
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
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        out = fn(*a, **k)
    return out, buf.getvalue()


def ref_knn(query, pool, k=32):
    """ddsp_prematch_dataset.py:1196-1206 driven exactly as the reference does."""
    idx, val = [], []
    for a in range(0, len(query), 20):
        d = ref_lib.fast_cosine_dist(query[a:a + 20], pool)
        tk = d.topk(k=k, dim=-1, largest=False)
        idx.append(tk.indices)
        val.append(tk.values)
    return torch.cat(idx), torch.cat(val)


def real_f0():
    sc = Path("/root/reference/sample_content")
    src = np.load(sc / "Danakil-voice_resampled_16000_cut_f0.npy").astype(np.float32)
    tgt = np.load(sc / "Tiken_lead_07_resampled_16000_cut_f0.npy").astype(np.float32)
    return src, tgt


def main():
    out = {}
    f0_src_full, f0_tgt_full = real_f0()

    # ---- K1 full matrix, both cdist forms
    a = synth.ar1_frames(30, seed=11)
    b = synth.ar1_frames(50, seed=12)
    out["dist_mm_30x50"] = ref_lib.fast_cosine_dist(torch.from_numpy(a), torch.from_numpy(b)).numpy()
    out["dist_direct_4x8"] = ref_lib.fast_cosine_dist(torch.from_numpy(a[:4]), torch.from_numpy(b[:8])).numpy()
    out["dist_mm_30x50_f64"] = ref_lib.fast_cosine_dist(torch.from_numpy(a).double(), torch.from_numpy(b).double()).numpy()

    # ---- K1+K2 kNN, AR(1) (WavLM-like) and randn, fp32 and fp64
    q = synth.ar1_frames(64, seed=1)
    p = synth.ar1_frames(700, seed=2)
    i32, v32 = ref_knn(torch.from_numpy(q), torch.from_numpy(p))
    i64, v64 = ref_knn(torch.from_numpy(q).double(), torch.from_numpy(p).double())
    out["knn_ar1_idx_f32"], out["knn_ar1_val_f32"] = i32.numpy(), v32.numpy()
    out["knn_ar1_idx_f64"], out["knn_ar1_val_f64"] = i64.numpy(), v64.numpy()
    qr = synth.randn_frames(50, seed=3)
    pr = synth.randn_frames(1500, seed=4)
    i32, v32 = ref_knn(torch.from_numpy(qr), torch.from_numpy(pr))
    out["knn_randn_idx_f32"], out["knn_randn_val_f32"] = i32.numpy(), v32.numpy()
    i64, v64 = ref_knn(torch.from_numpy(qr).double(), torch.from_numpy(pr).double())
    out["knn_randn_idx_f64"], out["knn_randn_val_f64"] = i64.numpy(), v64.numpy()

    # ---- K4 f0 re-rank on the real f0 tracks
    f0q = f0_src_full[1000:1064].copy()
    f0p = f0_tgt_full[500:1200].copy()
    out["f0_src_crop"], out["f0_tgt_crop"] = f0q, f0p
    nbrs = torch.from_numpy(out["knn_ar1_idx_f32"])
    prio = ref_pm.sort_by_f0_compatibility(torch.from_numpy(f0q), torch.from_numpy(f0p), nbrs)
    out["f0_prio_idx"] = prio.numpy()

    # ---- K5 greedy re-selection, no-f0 and f0 branches; resets force the sticky branch (D6)
    T5 = 150
    q5 = synth.ar1_frames(T5, seed=21, reset_every=60)
    p5 = synth.ar1_frames(600, seed=22)
    nb5, _ = ref_knn(torch.from_numpy(q5), torch.from_numpy(p5))
    f0q5 = f0_src_full[2000:2000 + T5].copy()
    f0p5 = f0_tgt_full[300:900].copy()
    out["k5_f0_src"], out["k5_f0_tgt"] = f0q5, f0p5
    out["k5_nbrs"] = nb5.numpy()
    for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        r, _ = quiet(ref_lib.knn_with_concat_cost, nb5[:, :4].clone(), torch.from_numpy(q5).to(dt),
                     torch.from_numpy(p5).to(dt), concat_weight=0.2)
        out[f"k5_nof0_{tag}"] = r.numpy()
        pr5 = ref_pm.sort_by_f0_compatibility(torch.from_numpy(f0q5), torch.from_numpy(f0p5), nb5)
        r, _ = quiet(ref_lib.knn_with_concat_cost, pr5[:, :4].clone(), torch.from_numpy(q5).to(dt),
                     torch.from_numpy(p5).to(dt), torch.from_numpy(f0q5), torch.from_numpy(f0p5), concat_weight=0.2)
        out[f"k5_f0_{tag}"] = r.numpy()
    out["k5_prio"] = pr5.numpy()

    # ---- K6 Adam weight fit (wavlm + extended), fp64 storage as on the real path and fp32
    T6 = 80
    p6 = synth.ar1_frames(400, seed=32)
    q6 = synth.ar1_frames(T6, seed=31)
    nb6, _ = ref_knn(torch.from_numpy(q6), torch.from_numpy(p6))
    idx6 = nb6[:, :4].clone()
    out["k6_idx"] = idx6.numpy()
    h6 = synth.harmonics_pool(400, seed=33)
    for dt, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        w, log = quiet(ref_pm.compute_wavlm_weight, idx6.clone(), torch.from_numpy(p6).to(dt), "sum_to_1_geq")
        its = [int(m) for m in re.findall(r"(?:^|\r|\n)(\d+) ", log)]
        out[f"k6_wavlm_w_{tag}"] = w.detach().numpy()
        out[f"k6_wavlm_last_t_{tag}"] = np.int64(max(its))
        w, log = quiet(ref_pm.compute_extended_weight, idx6.clone(), torch.from_numpy(h6).to(dt), "sum_to_1_geq", [1])
        its = [int(m) for m in re.findall(r"(?:^|\r|\n)(\d+) ", log)]
        out[f"k6_ext_w_{tag}"] = w.detach().numpy()
        out[f"k6_ext_last_t_{tag}"] = np.int64(max(its))

    # ---- K7 harmonic bank and K7' single sinusoid
    T7 = 24
    f07 = f0_src_full[1500:1500 + T7].copy()
    amp7 = synth.harmonics_pool(T7, seed=41)
    out["k7_f0"] = f07
    sig = ref_pm.get_bulk_dsp_choral(torch.from_numpy(f07)[None, :, None], torch.from_numpy(amp7)[None])
    out["k7_signal"] = sig.numpy()
    f07b = np.stack([f07, f0_tgt_full[700:700 + T7]])
    amp7b = np.stack([amp7, synth.harmonics_pool(T7, seed=42)])
    out["k7_f0_b2"] = f07b
    out["k7_signal_b2"] = ref_pm.get_bulk_dsp_choral(torch.from_numpy(f07b)[..., None], torch.from_numpy(amp7b)).numpy()
    out["k7_amp_up"] = ref_pm.upsample(torch.from_numpy(amp7)[None].transpose(1, 2), 320, mode="bicubic").transpose(1, 2).numpy()[:, ::37]
    # hifigan/ddsp_models_f0.py:344-352 with its local upsample (:99-102), executed op for op
    f0t = torch.from_numpy(f07)[None, :, None]
    pitch = torch.nn.functional.interpolate(f0t.permute(0, 2, 1), size=T7 * 320).permute(0, 2, 1)
    omega = torch.cumsum(pitch.double() / 16000, dim=1)
    omega = (2 * np.pi * (omega - torch.round(omega))).float()
    out["k7p_signal"] = torch.sin(omega).transpose(1, 2).numpy()

    # ---- a11 whole pipeline: match_at_inference_time with the pool builder stubbed out
    Tq, Np = 120, 400
    qf = synth.ar1_frames(Tq, seed=51, reset_every=50)
    pf = synth.ar1_frames(Np, seed=52)
    f0q = f0_src_full[100:100 + Tq].copy()
    f0p = f0_tgt_full[1300:1300 + Np].copy()
    hp = synth.harmonics_pool(Np, seed=53)
    out["pipe_f0_src"], out["pipe_f0_tgt"] = f0q, f0p

    def fake_pool(wav, *a, **k):
        if "src" in str(wav):
            feats, f0, n = torch.from_numpy(qf).double(), torch.from_numpy(f0q), Tq
            harm = torch.zeros(n, 49)
        else:
            feats, f0, n = torch.from_numpy(pf).double(), torch.from_numpy(f0p), Np
            harm = torch.from_numpy(hp)
        key = str(wav)
        return ({key: feats}, {key: feats}, {key: torch.zeros(n, 320)}, {key: torch.ones(n, 201)}, {key: f0}, {key: harm})

    ref_pm.get_complete_spk_pool = fake_pool
    for post_opt in ("no_post_opt", "post_opt_0.2"):
        res, _ = quiet(ref_pm.match_at_inference_time, Path("/x/src.wav"), Path("/x/ref.wav"), None, None, None,
                       device="cpu", prioritize_f0=True, ckpt_type="mix", src_dataset_path="/x",
                       tgt_dataset_path="/x", post_opt=post_opt)
        feats, harm, _, sf0 = res
        tag = post_opt.replace(".", "p")
        fe = feats["/x/src.wav"].detach().numpy()
        out[f"pipe_{tag}_feats_sub"] = fe[:, ::16]
        out[f"pipe_{tag}_feats_rowsum"] = fe.astype(np.float64).sum(1)
        out[f"pipe_{tag}_harm"] = harm["/x/src.wav"].detach().numpy()
        out[f"pipe_{tag}_f0"] = sf0["/x/src.wav"].numpy()

    np.savez_compressed(HERE / "reference_outputs.npz", **out)
    print({k: (v.shape, str(v.dtype)) for k, v in out.items()})
    print("bytes", os.path.getsize(HERE / "reference_outputs.npz"))


if __name__ == "__main__":
    main()
