"""Where a single utterance's post_opt_0.2 match goes: per-stage CUDA-event times (stages run back to back on
one stream) against the two-stream whole, for a 3001-frame and an 800-frame utterance."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from knn_svc_b200 import ops, synth
from knn_svc_b200 import ddsp_prematch_dataset as pm

dev = "cuda:0"


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return r, float(np.median(ts)), float(min(ts))


for T, Np in ((3001, 3001), (800, 30000)):
    q = torch.from_numpy(synth.ar1_frames(T, seed=1, reset_every=200)).to(dev)
    p = torch.from_numpy(synth.ar1_frames(min(Np, 6000), seed=2)).to(dev)
    if Np > 6000:
        g = torch.Generator(device=dev); g.manual_seed(1)
        p = p.repeat((Np + 5999) // 6000, 1)[:Np] + 0.3 * torch.randn((Np, 1024), device=dev, generator=g)
    f0q = torch.from_numpy(synth.f0_track(T, seed=3)); f0p = torch.from_numpy(synth.f0_track(Np, seed=4))
    hp = torch.from_numpy(synth.harmonics_pool(Np, seed=5))
    pool = pm.MatchingPool(p, p, f0p, hp, dev)
    out = {"T": T, "Np": Np}
    _, out["whole_two_streams_ms"], out["whole_min_ms"] = timed(lambda: pm.match_utterance(q, f0q, pool, post_opt="post_opt_0.2", ckpt_type="mix"))
    side = pm._side_stream
    pm._side_stream = lambda d: torch.cuda.current_stream(d)
    _, out["whole_one_stream_ms"], _ = timed(lambda: pm.match_utterance(q, f0q, pool, post_opt="post_opt_0.2", ckpt_type="mix"))
    pm._side_stream = side
    qp, out["prepare_ms"], _ = timed(lambda: ops.prepare_rows(q))
    (_, nn), out["search32_ms"], _ = timed(lambda: ops.knn_search(qp, pool.matching, 32))
    sf0, out["shift_f0_ms"], _ = timed(lambda: pm.shift_query_f0_batched([f0q], pool.log_f0_median))
    prio, out["f0_rerank_ms"], _ = timed(lambda: pm.sort_by_f0_compatibility(sf0, pool.f0_dev, nn))
    offs = [0, T]
    idx_h = prio[:, :4].contiguous(); idx_w = nn[:, :4].contiguous()
    ih, out["k5_f0_ms"], _ = timed(lambda: ops.concat_cost_reselect(idx_h, qp.rows, pool.matching.rows, sf0, pool.f0_dev, concat_weight=0.2, utt_offsets=offs))
    iw, out["k5_ms"], _ = timed(lambda: ops.concat_cost_reselect(idx_w, qp.rows, pool.matching.rows, concat_weight=0.2, utt_offsets=offs))
    hw, out["k6_ext_ms"], _ = timed(lambda: pm.compute_extended_weight(ih, pool.harmonics, "sum_to_1_geq", [1], utt_offsets=offs))
    w, out["k6_wavlm_ms"], _ = timed(lambda: pm.compute_wavlm_weight(iw, pool.synth, "sum_to_1_geq", utt_offsets=offs))
    _, out["mix_ms"], _ = timed(lambda: (ops.gather_mix(pool.synth, iw, w), ops.gather_mix(pool.harmonics, ih, hw)))
    print(json.dumps({k: (round(v, 3) if isinstance(v, float) else v) for k, v in out.items()}), flush=True)
