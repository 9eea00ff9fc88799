"""Per-kernel measurements of the non-GEMM rows of SURVEY §8 (K3-K7 + pre-pass), CUDA-event
timed, with achieved GB/s against the measured HBM peak.  Prints one JSON line per kernel.
Run under gpurun; the numbers go to profiles/."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from knn_svc_b200 import ops, synth

dev = "cuda:0"
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
HBM = peaks["hbm_gbs"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > L2


def timed(fn, reps=10, warm=3, flush_l2=True):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        if flush_l2:
            flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def report(name, ms, bytes_, frames, extra=None):
    gbs = bytes_ / (ms * 1e-3) / 1e9
    d = {"kernel": name, "ms": round(ms, 4), "frames_per_s": round(frames / (ms * 1e-3)), "algorithmic_GB": round(bytes_ / 1e9, 4),
         "achieved_GBs": round(gbs, 1), "hbm_peak_GBs": HBM, "frac_of_hbm_peak": round(gbs / HBM, 4)}
    if extra:
        d.update(extra)
    print(json.dumps(d), flush=True)


g = torch.Generator(device=dev); g.manual_seed(0)
D = 1024

# ---- prepare_rows: 2M rows
n = 2_000_000
x = torch.randn((n, D), device=dev, generator=g)
ms = timed(lambda: ops.prepare_rows(x, check=False), reps=5)
report("prepare_rows", ms, n * (D * 4 + D * 2 + 4), n)

# ---- K3 gather_mix: T=100k, K=4 from a 2M-row pool (features) and 49-wide harmonics
T = 100_000
idx = torch.randint(0, n, (T, 4), device=dev, generator=g)
w = torch.softmax(torch.randn((T, 4), device=dev, generator=g), 1)
ms = timed(lambda: ops.gather_mix(x, idx, w))
report("gather_mix D=1024 K=4", ms, T * (4 * D * 4 + D * 4 + 4 * 12), T)
hp = torch.rand((n, 49), device=dev, generator=g)
ms = timed(lambda: ops.gather_mix(hp, idx, w))
report("gather_mix D=49 K=4 (harmonics)", ms, T * (4 * 49 * 4 + 49 * 4 + 4 * 12), T)
del x

# ---- K4 f0 re-rank: T=100k, k=32
f0p = torch.rand(n, device=dev, generator=g) * 500 + 80
f0q = torch.rand(T, device=dev, generator=g) * 500 + 80
idx32 = torch.randint(0, n, (T, 32), device=dev, generator=g)
ms = timed(lambda: ops.f0_rerank(f0q, f0p, idx32))
report("f0_rerank k=32", ms, T * (32 * (8 + 4) + 4 + 32 * 8), T)

# ---- K5 greedy re-selection: cfg-2 shape (one utterance of 3001 frames, pool 3001) and a cfg-5-like batch
for (n_utt, t_utt, n_pool) in ((1, 3001, 3001), (64, 800, 30000), (512, 800, 30000)):
    Tq = n_utt * t_utt
    q = torch.from_numpy(synth.ar1_frames(min(Tq, 4000), seed=1)).to(dev).repeat((Tq + 3999) // 4000, 1)[:Tq].contiguous()
    p = torch.from_numpy(synth.ar1_frames(min(n_pool, 4000), seed=2)).to(dev).repeat((n_pool + 3999) // 4000, 1)[:n_pool].contiguous()
    p = p + 0.01 * torch.randn(p.shape, device=dev, generator=g)
    _, nb = ops.knn_search(ops.prepare_rows(q), ops.prepare_rows(p), 4)
    offs = [i * t_utt for i in range(n_utt + 1)]
    f0s = torch.from_numpy(synth.f0_track(Tq, seed=3)).to(dev)
    f0t = torch.from_numpy(synth.f0_track(n_pool, seed=4)).to(dev)
    ms = timed(lambda: ops.concat_cost_reselect(nb, q, p, utt_offsets=offs), reps=3, warm=1)
    report(f"concat_cost (no f0) utt={n_utt}x{t_utt}", ms, Tq * 9 * D * 4, Tq, {"us_per_frame_per_utt": round(ms * 1e3 / t_utt, 3)})
    ms = timed(lambda: ops.concat_cost_reselect(nb, q, p, f0s, f0t, utt_offsets=offs), reps=3, warm=1)
    report(f"concat_cost (f0) utt={n_utt}x{t_utt}", ms, Tq * 9 * D * 4, Tq, {"us_per_frame_per_utt": round(ms * 1e3 / t_utt, 3)})
    if n_utt == 1:
        # ---- K6 weight fit on the same utterance (features, then harmonics)
        out, info = ops.weight_fit(nb, p, 0.1, return_info=True)
        ms = timed(lambda: ops.weight_fit(nb, p, 0.1), reps=3, warm=1, flush_l2=False)
        report("weight_fit wavlm T=3001", ms, Tq * 3 * 4 * D * 4, Tq, {"adam_iterations": int(info[0].item()), "best_loss": float(info[1].item())})
        hpool = torch.from_numpy(synth.harmonics_pool(n_pool, seed=5)).to(dev)
        out, info = ops.weight_fit(nb, hpool, 1000.0, return_info=True)
        ms = timed(lambda: ops.weight_fit(nb, hpool, 1000.0), reps=3, warm=1, flush_l2=False)
        report("weight_fit harmonics T=3001", ms, Tq * 3 * 4 * 49 * 4, Tq, {"adam_iterations": int(info[0].item())})

# ---- K7 harmonic bank: T=3001 (cfg 1/2: 960 320 samples) and a batch of 64 such utterances
for B in (1, 64):
    f0 = torch.from_numpy(np.stack([synth.f0_track(3001, seed=10 + b) for b in range(min(B, 4))])).to(dev).repeat((B + 3) // 4, 1)[:B].contiguous()
    amp = torch.from_numpy(synth.harmonics_pool(3001, seed=6)).to(dev)[None].repeat(B, 1, 1).contiguous()
    ms = timed(lambda: ops.harmonic_bank(f0, amp))
    samples = B * 3001 * 320
    report(f"harmonic_bank H=49 B={B}", ms, B * 3001 * (4 + 49 * 4 + 320 * 4), B * 3001,
           {"Gsamples_per_s": round(samples / (ms * 1e-3) / 1e9, 3), "sinf_per_s": round(samples * 49 / (ms * 1e-3) / 1e12, 3)})
    ms = timed(lambda: ops.harmonic_bank(f0, None))
    report(f"f0_sinusoid B={B}", ms, B * 3001 * (4 + 320 * 4), B * 3001)

# ---- pool-builder ops (a 10-minute recording: 30k frames) and the offline prematch
Tl, L = 30_000, 25
layers = torch.randn((L, Tl, D), device=dev, generator=g)
wa = np.random.RandomState(0).rand(L); wb = np.zeros(L); wb[6] = 1.0
ms = timed(lambda: ops.layer_mix(layers, wa, wb), reps=5)
report("layer_mix L=25 (two mixes, one pass)", ms, Tl * (L * D * 4 + 2 * D * 4), Tl)
del layers
audio = torch.randn(320 * Tl + 80, device=dev, generator=g) * 0.1
ms = timed(lambda: ops.stft_magnitude(audio, Tl))
report("stft_magnitude n_fft=400 hop=320", ms, Tl * (320 * 4 + 200 * 4), Tl, {"GFMA_per_s": round(Tl * 400 * 200 * 2 / (ms * 1e-3) / 1e9, 1)})
spec = ops.stft_magnitude(audio, Tl)
f0u = torch.from_numpy(synth.f0_track(Tl, seed=7)).to(dev)
ms = timed(lambda: ops.harmonic_amplitudes(spec, f0u))
report("harmonic_amplitudes H=49", ms, Tl * (200 * 4 + 4 + 49 * 4), Tl)
spec_pool = torch.rand((2_000_000, 200), device=dev, generator=g)
ms = timed(lambda: ops.row_l1(spec_pool))
report("row_l1 S=200", ms, 2_000_000 * (200 * 4 + 4), 2_000_000)
del spec_pool
# masked self-search: one speaker's pool (256k frames = 85 min) against itself, k=32, utterances of 512 frames
ns = 262_144
xs = ops.prepare_rows(torch.randn((ns, D), device=dev, generator=g), check=False)
starts = torch.arange(ns // 512, device=dev) * 512
lens = torch.full((ns // 512,), 512, device=dev)
mlo, mhi = torch.repeat_interleave(starts, lens), torch.repeat_interleave(starts + 512, lens)
ms_m = timed(lambda: ops.knn_search(xs, xs, 32, mask_lo=mlo, mask_hi=mhi), reps=3, warm=1, flush_l2=False)
ms_u = timed(lambda: ops.knn_search(xs, xs, 32), reps=3, warm=1, flush_l2=False)
print(json.dumps({"kernel": "prematch self-search 262144^2 k=32 (filter+rescore)", "ms_masked": round(ms_m, 2), "ms_unmasked": round(ms_u, 2),
                  "frames_per_s_masked": round(ns / (ms_m * 1e-3)), "TFLOPs_masked": round(2.0 * ns * ns * D / (ms_m * 1e-3) / 1e12, 1),
                  "TFLOPs_unmasked": round(2.0 * ns * ns * D / (ms_u * 1e-3) / 1e12, 1)}), flush=True)
