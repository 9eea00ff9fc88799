import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from knn_svc_b200 import ops, synth
from knn_svc_b200 import ddsp_prematch_dataset as pm
dev = "cuda:0"
g = torch.Generator(device=dev); g.manual_seed(0)
mean = 3.0 * torch.randn(1024, device=dev, generator=torch.Generator(device=dev).manual_seed(12345))
pf = torch.randn((30000, 1024), device=dev, generator=g) + mean
pool = pm.MatchingPool(pf, pf, torch.from_numpy(synth.f0_track(30000, seed=0)),
                       torch.from_numpy(synth.harmonics_pool(2000, seed=5)).repeat(16, 1)[:30000], dev)
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return r, float(np.median(ts))
for T in (300, 800, 1400):
    q = torch.randn((T, 1024), device=dev, generator=g) * 0.6 + mean
    f0q = torch.from_numpy(synth.f0_track(T, seed=1000))
    out = {"T": T}
    _, out["whole_ms"] = timed(lambda: pm.match_utterance(q, f0q, pool, post_opt="post_opt_0.2", ckpt_type="mix"))
    qp = ops.prepare_rows(q)
    (_, nn, st), out["search32_ms"] = timed(lambda: ops.knn_search(qp, pool.matching, 32, return_stats=True))
    out["stats"] = st.tolist()
    sf0 = pm.shift_query_f0_batched([f0q], pool.log_f0_median)
    prio = pm.sort_by_f0_compatibility(sf0, pool.f0_dev, nn)
    offs = [0, T]
    idx_h = prio[:, :4].contiguous(); idx_w = nn[:, :4].contiguous()
    ih, out["k5_f0_ms"] = timed(lambda: ops.concat_cost_reselect(idx_h, qp.rows, pool.matching.rows, sf0, pool.f0_dev, concat_weight=0.2, utt_offsets=offs))
    iw, out["k5_ms"] = timed(lambda: ops.concat_cost_reselect(idx_w, qp.rows, pool.matching.rows, concat_weight=0.2, utt_offsets=offs))
    (hw, hinfo), out["k6_ext_ms"] = timed(lambda: ops.weight_fit(ih, pool.harmonics, 1000.0, return_info=True, utt_offsets=offs))
    (w, winfo), out["k6_wavlm_ms"] = timed(lambda: ops.weight_fit(iw, pool.synth, 0.1, return_info=True, utt_offsets=offs))
    out["k6_ext_iters"] = float(hinfo.flatten()[0]); out["k6_wavlm_iters"] = float(winfo.flatten()[0])
    print(json.dumps(out), flush=True)
