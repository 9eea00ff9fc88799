"""knn_search at the shapes of BASELINE cfg 2/3/5 and the offline prematch: total time (CUDA events)
and the tcgen05 filter's share (knnsvc_filter_timing); the rest is re-scoring + bookkeeping.
Also times cuBLAS fp16 sustained (the filter's operand type) beside MEASURED_PEAKS' bf16 figure."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from knn_svc_b200 import _lib, ops, synth

dev = "cuda:0"
lib = _lib.load()
g = torch.Generator(device=dev); g.manual_seed(0)


def run(name, q, p, k, reps=5, **kw):
    qp, pp = ops.prepare_rows(q, check=False), (ops.prepare_rows(p, check=False) if p is not None else None)
    pp = qp if pp is None else pp
    for _ in range(2):
        d, i, st = ops.knn_search(qp, pp, k, return_stats=True, **kw)
    torch.cuda.synchronize()
    lib.knnsvc_filter_timing(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.knn_search(qp, pp, k, **kw)
    e1.record(); torch.cuda.synchronize()
    buf = (ctypes.c_float * 256)()
    n = lib.knnsvc_filter_timing_collect(ctypes.cast(buf, ctypes.c_void_p), 256)
    lib.knnsvc_filter_timing(0)
    tot = e0.elapsed_time(e1) / reps
    filt = sum(buf[j] for j in range(n)) / max(n, 1)
    T, NP = qp.n, pp.n
    print(json.dumps({"case": name, "T": T, "Np": NP, "k": k, "total_ms": round(tot, 3), "filter_ms": round(filt, 3),
                      "rest_ms": round(tot - filt, 3), "filter_TFLOPs": round(2.0 * T * NP * 1024 / filt / 1e9, 1),
                      "qframes_per_s": round(T / tot * 1e3), "logged_per_row": round(int(st[1]) / T, 1),
                      "survivors_per_row": round(int(st[2]) / T, 1), "n_seg": int(st[3]), "flagged": int(st[0])}), flush=True)


def ar1(n, seed):
    base = torch.from_numpy(synth.ar1_frames(min(n, 6000), seed=seed)).to(dev)
    x = base.repeat((n + len(base) - 1) // len(base), 1)[:n].contiguous()
    return x + 0.05 * torch.randn(x.shape, device=dev, generator=g)


for gen in ("randn", "ar1"):
    mk = (lambda n, s: torch.randn((n, 1024), device=dev, generator=g)) if gen == "randn" else ar1
    for (T, NP) in ((3000, 30000), (100_000, 30_000)):
        q, p = mk(T, 1), mk(NP, 2)
        for k in (4, 32):
            run(f"{gen} cfg3/5-shape", q, p, k)
x = torch.randn((262144, 1024), device=dev, generator=g)
starts = torch.arange(512, device=dev) * 512
lens = torch.full((512,), 512, device=dev)
run("randn prematch self-search (masked)", x, None, 32, reps=2, mask_lo=torch.repeat_interleave(starts, lens),
    mask_hi=torch.repeat_interleave(starts + 512, lens))
del x
# cuBLAS fp16 vs bf16, sustained 3 s each (same data distribution as the filter's operands)
for dt in (torch.float16, torch.bfloat16):
    a = torch.randn((8192, 8192), device=dev, generator=g).to(dt); b = torch.randn((8192, 8192), device=dev, generator=g).to(dt)
    for _ in range(20):
        a @ b
    torch.cuda.synchronize()
    import time
    t0 = time.time(); n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < 3.0:
        for _ in range(50):
            a @ b
        n += 50
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    print(json.dumps({"case": f"cuBLAS {dt} 8192^3 sustained", "TFLOPs": round(2 * 8192 ** 3 * n / e0.elapsed_time(e1) / 1e9, 1)}), flush=True)
