"""BASELINE cfg 5 shape: dataset->dataset conversion, post_opt_0.2, synthetic features.
(utterance, target speaker) pairs are independent; each target pool (30k frames) serves a batch
of source utterances (lengths ~U(150,1500) frames).  Measures matcher pairs/s and query frames/s
on one GPU (pairs are dealt round-robin over GPUs in the 8-GPU run: no communication).
    python tools/bench_cfg5.py [--utts 128] [--pools 2]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from knn_svc_b200 import ops, synth
from knn_svc_b200 import ddsp_prematch_dataset as pm

ap = argparse.ArgumentParser()
ap.add_argument("--utts", type=int, default=128)
ap.add_argument("--pools", type=int, default=2)
ap.add_argument("--pool-frames", type=int, default=30000)
ap.add_argument("--post-opt", default="post_opt_0.2")
ap.add_argument("--single", action="store_true", help="also time the per-utterance path on a subset")
args = ap.parse_args()
dev = "cuda:0"
g = torch.Generator(device=dev); g.manual_seed(0)
rs = np.random.RandomState(0)


lens = rs.randint(150, 1501, size=args.utts).tolist()
pools = []
for p in range(args.pools):
    pf = torch.randn((args.pool_frames, 1024), device=dev, generator=g) + 3.0 * torch.randn(1024, device=dev, generator=torch.Generator(device=dev).manual_seed(12345))
    pools.append(pm.MatchingPool(pf, pf, torch.from_numpy(synth.f0_track(args.pool_frames, seed=p)),
                                 torch.from_numpy(synth.harmonics_pool(2000, seed=p + 5)).repeat(args.pool_frames // 2000 + 1, 1)[:args.pool_frames], dev))
mean = 3.0 * torch.randn(1024, device=dev, generator=torch.Generator(device=dev).manual_seed(12345))
qs = [torch.randn((n, 1024), device=dev, generator=g) * 0.6 + mean for n in lens]
f0s = [torch.from_numpy(synth.f0_track(n, seed=1000 + i)) for i, n in enumerate(lens)]
total_frames = sum(lens) * args.pools


def run_batched():
    out = []
    for pool in pools:
        out.append(pm.match_utterances(qs, f0s, pool, post_opt=args.post_opt, ckpt_type="mix", prioritize_f0=True))
    return out


run_batched(); torch.cuda.synchronize()
dts = []
for _ in range(7):
    t0 = time.perf_counter(); run_batched(); torch.cuda.synchronize(); dts.append(time.perf_counter() - t0)
dt = float(np.median(dts))
line = {"workload": f"cfg5 shape: {args.utts} utterances x {args.pools} target pools of {args.pool_frames} frames, {args.post_opt}",
        "pairs": args.utts * args.pools, "query_frames": total_frames, "seconds": dt,
        "pairs_per_s": args.utts * args.pools / dt, "query_frames_per_s": total_frames / dt, "path": "match_utterances (batched)",
        "timing": "median of 7 passes, wall clock incl. host orchestration", "min_s": min(dts), "max_s": max(dts)}
print(json.dumps(line), flush=True)
if args.single:
    n = min(16, args.utts)
    for q, f in zip(qs[:n], f0s[:n]):
        pm.match_utterance(q, f, pools[0], post_opt=args.post_opt, ckpt_type="mix", prioritize_f0=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for q, f in zip(qs[:n], f0s[:n]):
        pm.match_utterance(q, f, pools[0], post_opt=args.post_opt, ckpt_type="mix", prioritize_f0=True)
    torch.cuda.synchronize(); dt1 = time.perf_counter() - t0
    print(json.dumps({"path": "match_utterance (one by one)", "pairs": n, "seconds": dt1, "pairs_per_s": n / dt1,
                      "query_frames_per_s": sum(lens[:n]) / dt1}), flush=True)
