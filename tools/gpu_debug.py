"""Bring-up diagnostics for the CUDA path (run under gpurun).  Not part of the product."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from knn_svc_b200 import ops, synth
from oracle import matcher_oracle as orc

DEV = "cuda:0"
torch.cuda.set_device(0)
print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0), flush=True)


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(DEV)


def knn_case(T, Np, D, k, gen="randn"):
    q = (synth.randn_frames(T, d=D, seed=1) if gen == "randn" else synth.ar1_frames(T, d=D, seed=1))
    p = (synth.randn_frames(Np, d=D, seed=2) if gen == "randn" else synth.ar1_frames(Np, d=D, seed=2))
    qp, pp = ops.prepare_rows(t(q)), ops.prepare_rows(t(p))
    torch.cuda.synchronize()
    t0 = time.time()
    dist, idx, stats = ops.knn_search(qp, pp, k, return_stats=True)
    torch.cuda.synchronize()
    dt = time.time() - t0
    e_dist, e_idx = ops.knn_exact(qp, pp, k)
    torch.cuda.synchronize()
    same = (idx == e_idx).float().mean().item()
    md = (dist - e_dist).abs().max().item()
    print(f"knn {gen} T={T} Np={Np} D={D} k={k}: idx agree {same:.4f} max|dd| {md:.2e} stats {stats.tolist()} {dt*1e3:.1f} ms",
          flush=True)
    if T * Np <= 4e6:
        o_idx, o_val = orc.knn(q, p, k)
        print("   vs oracle: exact-kernel idx agree", (e_idx.cpu().numpy() == o_idx).mean(),
              "filter idx agree", (idx.cpu().numpy() == o_idx).mean(),
              "max|d-oracle|", np.abs(dist.cpu().numpy() - o_val).max(), flush=True)
    return same


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "knn"):
        knn_case(64, 700, 1024, 4)
        knn_case(64, 700, 1024, 32, "ar1")
        knn_case(300, 5000, 1024, 32)
        knn_case(129, 513, 192, 8)
        knn_case(3000, 30000, 1024, 32)
    if which in ("all", "perf"):
        for (T, Np) in ((3000, 30000), (16384, 262144), (32768, 1048576)):
            g = torch.Generator(device=DEV); g.manual_seed(0)
            q = torch.randn((T, 1024), device=DEV, generator=g)
            p = torch.randn((Np, 1024), device=DEV, generator=g)
            qp, pp = ops.prepare_rows(q), ops.prepare_rows(p)
            for k in (4, 32):
                ops.knn_search(qp, pp, k)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                n = 3
                for _ in range(n):
                    d, i, st = ops.knn_search(qp, pp, k, return_stats=True)
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / n
                fl = 2.0 * T * Np * 1024
                print(f"perf T={T} Np={Np} k={k}: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s  {T/ms*1e3:.0f} qf/s stats {st.tolist()}",
                      flush=True)
