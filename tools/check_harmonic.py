"""GPU check + timing of the harmonic bank (K7) against the CPU oracle over edge shapes.
Run under gpurun: python tools/check_harmonic.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from knn_svc_b200 import ops, synth
from oracle import matcher_oracle as orc   # checker only

dev = "cuda:0"
worst = 0.0
cases = [(3001, 49, 320, 3), (257, 49, 320, 5), (64, 1, 320, 6), (100, 7, 320, 7), (90, 50, 320, 8), (50, 64, 320, 9),
         (40, 65, 320, 10), (33, 49, 321, 11), (35, 49, 5, 12), (1, 49, 320, 13), (2, 49, 320, 14), (3, 3, 2, 15)]
for T, H, hop, seed in cases:
    f0 = synth.f0_track(T, seed=seed)
    if seed % 2 == 0:
        f0[: T // 3] *= 6.0           # push harmonics above Nyquist
    amp = synth.harmonics_pool(T, h=H, seed=seed + 100)
    got = ops.harmonic_bank(torch.from_numpy(f0).to(dev)[None], torch.from_numpy(amp).to(dev)[None], 16000, hop).cpu().numpy()
    want = orc.get_bulk_dsp_choral(f0[None, :, None], amp[None], 16000, hop)[..., 0]
    err = np.abs(got - want).max() / max(np.abs(want).max(), 1e-30)
    worst = max(worst, err)
    print(json.dumps({"T": T, "H": H, "hop": hop, "rel_err": float(err), "ok": bool(err <= 1e-4)}), flush=True)
assert worst <= 1e-4, worst

def timed(fn, reps=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))

for B in (1, 64):
    T = 3001
    f0 = torch.from_numpy(np.stack([synth.f0_track(T, seed=s) for s in range(B)])).to(dev)
    amp = torch.from_numpy(np.stack([synth.harmonics_pool(T, seed=s) for s in range(min(B, 4))] * (B // min(B, 4)))).to(dev)
    ms = timed(lambda: ops.harmonic_bank(f0, amp))
    print(json.dumps({"kernel": f"harmonic_bank H=49 B={B} T={T}", "ms": ms, "frames_per_s": B * T / ms * 1e3,
                      "Gsamples_per_s": B * T * 320 / ms / 1e6, "algorithmic_GBs": B * T * 1480 / ms / 1e6}), flush=True)
