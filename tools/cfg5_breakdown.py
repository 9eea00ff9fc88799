"""Where a cfg-5 batch goes: per-stage times (stages back to back on one stream, wall clock with a
synchronize after each) against the two-stream whole, for one target pool of 30k frames and a batch
of utterances with U(150,1500) frames (the per-rank batch of the 8-GPU run is ~220, of the 1-GPU run
up to 1024).  Also the search alone with its filter share and candidate statistics."""
import argparse, ctypes, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from knn_svc_b200 import _lib, ops, synth
from knn_svc_b200 import ddsp_prematch_dataset as pm

ap = argparse.ArgumentParser()
ap.add_argument("--utts", default="220,880")
args = ap.parse_args()
dev = "cuda:0"
lib = _lib.load()


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return r, float(np.median(ts))


rows = synth.ar1_frames_device(30000, 1024, seed=900, device=dev, seg_len=500)
f0p = torch.from_numpy(synth.f0_track(30000, seed=901))
harm = torch.from_numpy(synth.harmonics_pool(2000, seed=902)).repeat(16, 1)[:30000]
pool = pm.MatchingPool(rows, rows, f0p, harm, dev)
for n_utt in [int(v) for v in args.utts.split(",")]:
    rs = np.random.RandomState(n_utt)
    lens = sorted((int(v) for v in rs.randint(150, 1501, size=n_utt)), reverse=True)
    qs = [synth.ar1_frames_device(n, 1024, seed=5000 + i, device=dev, seg_len=200) for i, n in enumerate(lens)]
    f0s = [torch.from_numpy(synth.f0_track(n, seed=7000 + i)) for i, n in enumerate(lens)]
    offs = [0]
    for n in lens:
        offs.append(offs[-1] + n)
    out = {"utterances": n_utt, "frames": offs[-1]}
    _, out["whole_two_streams_ms"] = timed(lambda: pm.match_utterances(qs, f0s, pool, post_opt="post_opt_0.2", ckpt_type="mix"))
    qcat, out["concat_ms"] = timed(lambda: torch.concat(qs, dim=0))
    qp, out["prepare_ms"] = timed(lambda: ops.prepare_rows(qcat))
    lib.knnsvc_filter_timing(1)
    (_, nn, st), out["search32_ms"] = timed(lambda: ops.knn_search(qp, pool.matching, 32, return_stats=True))
    buf = (ctypes.c_float * 256)()
    n = lib.knnsvc_filter_timing_collect(ctypes.cast(buf, ctypes.c_void_p), 256)
    lib.knnsvc_filter_timing(0)
    out["search32_filter_ms"] = sum(buf[j] for j in range(n)) / max(n, 1)
    out["logged_per_row"], out["survivors_per_row"], out["fp64_scored_per_row"] = int(st[1]) / offs[-1], int(st[2]) / offs[-1], int(st[7]) / offs[-1]
    sf0, out["shift_f0_ms"] = timed(lambda: pm.shift_query_f0_batched(f0s, pool.log_f0_median))
    prio, out["f0_rerank_ms"] = timed(lambda: pm.sort_by_f0_compatibility(sf0, pool.f0_dev, nn))
    idx_h = prio[:, :4].contiguous(); idx_w = nn[:, :4].contiguous()
    ih, out["k5_f0_ms"] = timed(lambda: ops.concat_cost_reselect(idx_h, qp.rows, pool.matching.rows, sf0, pool.f0_dev, concat_weight=0.2, utt_offsets=offs))
    iw, out["k5_ms"] = timed(lambda: ops.concat_cost_reselect(idx_w, qp.rows, pool.matching.rows, concat_weight=0.2, utt_offsets=offs))
    hw, out["k6_ext_ms"] = timed(lambda: pm.compute_extended_weight(ih, pool.harmonics, "sum_to_1_geq", [1], utt_offsets=offs))
    w, out["k6_wavlm_ms"] = timed(lambda: pm.compute_wavlm_weight(iw, pool.synth, "sum_to_1_geq", utt_offsets=offs))
    _, out["mix_ms"] = timed(lambda: (ops.gather_mix(pool.synth, iw, w), ops.gather_mix(pool.harmonics, ih, hw)))
    out["sum_of_stages_ms"] = sum(v for k, v in out.items() if k.endswith("_ms") and k not in ("whole_two_streams_ms", "search32_filter_ms"))
    print(json.dumps({k: (round(v, 3) if isinstance(v, float) else v) for k, v in out.items()}), flush=True)
    del qs, qcat, qp
