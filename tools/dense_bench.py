"""Search on dense (WavLM-like) data: filter time vs whole-search time, candidate statistics.

Cases:  "tiled" = the round-1 A/B generator (6000 AR(1) rows tiled to size
+ 0.05 noise: ~170 near-duplicates of every pool row at cosine distance ~2.5e-4 — the 20k x 1M k=4
case the judge quoted at 4.2x), "G" = generator G proper (independent AR(1) runs + shared mean).
    python tools/dense_bench.py [--cases tiled:20000:1000000:4,...] [--reps 3] [--check]"""
import argparse, ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from knn_svc_b200 import _lib, ops, synth

ap = argparse.ArgumentParser()
ap.add_argument("--cases", default="tiled:20000:1000000:4,tiled:20000:1000000:32,G:20000:1000000:4,G:20000:1000000:32,"
                                   "tiled:100000:30000:4,tiled:100000:30000:32,tiled:3000:30000:32,randn:100000:1000000:4")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--check", action="store_true", help="compare a sample of rows with the exact CUDA-core kernel")
ap.add_argument("--out", default="")
args = ap.parse_args()
dev = "cuda:0"
lib = _lib.load()
g = torch.Generator(device=dev); g.manual_seed(0)


def tiled(n, seed):
    base = torch.from_numpy(synth.ar1_frames(min(n, 6000), seed=seed)).to(dev)
    x = base.repeat((n + len(base) - 1) // len(base), 1)[:n].contiguous()
    return x + 0.05 * torch.randn(x.shape, device=dev, generator=g)


def make(gen, n, seed, seg):
    if gen == "tiled":
        return tiled(n, seed)
    if gen == "G":
        return synth.ar1_frames_device(n, 1024, seed=seed, device=dev, seg_len=seg)
    return torch.randn((n, 1024), device=dev, generator=g)


if os.environ.get("REFINE_MIN"):
    lib.knnsvc_set_option(b"refine_min_candidates", int(os.environ["REFINE_MIN"]))
lines = []
for case in args.cases.split(","):
    gen, T, NP, k = case.split(":"); T, NP, k = int(T), int(NP), int(k)
    g.manual_seed(0)
    q, p = make(gen, T, 1, 200), make(gen, NP, 2, 500)
    qp, pp = ops.prepare_rows(q, check=False), ops.prepare_rows(p, check=False)
    for _ in range(2):
        d, i, st = ops.knn_search(qp, pp, k, return_stats=True)
    torch.cuda.synchronize()
    lib.knnsvc_filter_timing(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        ops.knn_search(qp, pp, k)
    e1.record(); torch.cuda.synchronize()
    buf = (ctypes.c_float * 256)()
    n = lib.knnsvc_filter_timing_collect(ctypes.cast(buf, ctypes.c_void_p), 256)
    lib.knnsvc_filter_timing(0)
    tot = e0.elapsed_time(e1) / args.reps
    filt = sum(buf[j] for j in range(n)) / max(n, 1)
    line = {"case": case, "total_ms": round(tot, 3), "filter_ms": round(filt, 3), "rest_ms": round(tot - filt, 3),
            "search_over_filter": round(tot / filt, 3), "filter_TFLOPs": round(2.0 * T * NP * 1024 / filt / 1e9, 1),
            "logged_per_row": round(int(st[1]) / T, 1), "survivors_per_row": round(int(st[2]) / T, 1),
            "flagged_rows": int(st[0]), "n_seg": int(st[3]), "cap": int(st[6])}
    if args.check:
        rows = torch.linspace(0, T - 1, 48, device=dev).long().unique()
        qs = ops.prepare_rows(q[rows].contiguous(), check=False)
        de, ie = ops.knn_exact(qs, pp, k)
        line["sample_rows_equal_exact_kernel"] = bool((ie == i[rows]).all()) and bool((de == d[rows]).all())
    print(json.dumps(line), flush=True)
    lines.append(line)
    del q, p, qp, pp
if args.out:
    with open(args.out, "w") as f:
        for l in lines:
            f.write(json.dumps(l) + "\n")
