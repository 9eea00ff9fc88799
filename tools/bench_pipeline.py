"""Whole-utterance matcher timings at the BASELINE cfg 1/2/3 shapes (device-resident inputs)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from knn_svc_b200 import ops, synth
from knn_svc_b200 import ddsp_prematch_dataset as pm

dev = "cuda:0"
def run(T, Np, post_opt, reps=5):
    q = torch.from_numpy(synth.ar1_frames(T, seed=1, reset_every=200)).to(dev)
    p = torch.from_numpy(synth.ar1_frames(min(Np, 6000), seed=2)).to(dev)
    if Np > 6000:
        g = torch.Generator(device=dev); g.manual_seed(1)
        p = p.repeat((Np + 5999) // 6000, 1)[:Np] + 0.3 * torch.randn((Np, 1024), device=dev, generator=g)
    f0q = torch.from_numpy(synth.f0_track(T, seed=3)); f0p = torch.from_numpy(synth.f0_track(Np, seed=4))
    hp = torch.from_numpy(synth.harmonics_pool(Np, seed=5))
    pool = pm.MatchingPool(p, p, f0p, hp, dev)
    def step():
        r = pm.match_utterance(q, f0q, pool, post_opt=post_opt, ckpt_type="mix")
        sig = pm.get_bulk_dsp_choral(r["shifted_f0"].to(dev)[None, :, None], r["harmonics"][None])
        return r, sig
    step(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); step(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    ms = float(np.median(ts)) * 1e3
    print(json.dumps({"pipeline": f"match_utterance+harmonic_bank T={T} Np={Np} {post_opt} (mix, prioritize_f0)",
                      "ms": round(ms, 3), "query_frames_per_s": round(T / ms * 1e3)}), flush=True)

run(3001, 3001, "no_post_opt")
run(3001, 3001, "post_opt_0.2")
run(3000, 30000, "no_post_opt")
run(3000, 30000, "post_opt_0.2")
