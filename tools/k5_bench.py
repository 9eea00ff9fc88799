"""K5 (knn_with_concat_cost) kernels side by side: general (post.cu), one-CTA staged, cluster of 8 CTAs per
utterance.  Prints one JSON line per (kernel, case): ms per launch and us per frame and utterance.
    python tools/k5_bench.py [--out profiles/...jsonl]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from knn_svc_b200 import _lib, ops

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="")
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
dev = "cuda:0"
lib = _lib.load()
KERNELS = {"general": (0, 0), "staged": (1, 0), "cluster": (1, 1)}
g = torch.Generator(device=dev); g.manual_seed(0)
lines = []
# (frames per utterance, utterances, pool frames)
for T, n_utt, Np in ((3001, 1, 3001), (3001, 1, 30000), (3001, 4, 30000), (3001, 16, 30000), (300, 16, 30000)):
    q = torch.randn((T * n_utt, 1024), device=dev, generator=g)
    p = torch.randn((Np, 1024), device=dev, generator=g)
    # candidates that mostly continue the previous frame's (as on speech), so the +1 rows matter
    base = torch.randint(0, max(1, Np - T - 8), (n_utt, 1), device=dev, generator=g) + torch.arange(T, device=dev)[None]
    idx = (base.reshape(-1, 1) + torch.randint(-2, 3, (T * n_utt, 4), device=dev, generator=g)).clamp_(0, Np - 1)
    f0q = torch.rand(T * n_utt, device=dev, generator=g) * 300 + 100
    f0p = torch.rand(Np, device=dev, generator=g) * 300 + 100
    offs = [i * T for i in range(n_utt + 1)]
    outs = {}
    for use_f0 in (False, True):
        for name, (st, cl) in KERNELS.items():
            lib.knnsvc_set_option(b"concat_staged", st); lib.knnsvc_set_option(b"concat_cluster", cl)
            a = (idx, q, p) + ((f0q, f0p) if use_f0 else (None, None))
            for _ in range(2):
                out = ops.concat_cost_reselect(*a, concat_weight=0.2, utt_offsets=offs)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                ops.concat_cost_reselect(*a, concat_weight=0.2, utt_offsets=offs)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            outs[(use_f0, name)] = out
            line = {"kernel": name, "frames": T, "utterances": n_utt, "pool": Np, "f0": use_f0, "ms": round(ms, 3),
                    "us_per_frame": round(ms * 1e3 / T, 3),
                    "equals_staged": bool(torch.equal(out, outs[(use_f0, "staged")])) if name == "cluster" else None}
            print(json.dumps(line), flush=True)
            lines.append(line)
lib.knnsvc_set_option(b"concat_staged", 1); lib.knnsvc_set_option(b"concat_cluster", 1)
if args.out:
    with open(args.out, "w") as f:
        for l in lines:
            f.write(json.dumps(l) + "\n")
