"""A/B of static (filter_flags=5) vs dynamic (1) unit scheduling on small search shapes, alternating in one process."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from knn_svc_b200 import _lib, ops

lib = _lib.load()
dev = "cuda:0"
g = torch.Generator(device=dev); g.manual_seed(0)
shapes = [(3000, 30000, 4), (3000, 30000, 32), (3001, 3001, 32), (1500, 30000, 4), (20000, 30000, 4), (3000, 1000000, 4)]
for (T, NP, K) in shapes:
    q = torch.randn((T, 1024), device=dev, generator=g); p = torch.randn((NP, 1024), device=dev, generator=g)
    qp, pp = ops.prepare_rows(q, check=False), ops.prepare_rows(p, check=False)
    for rep in range(3):
        for flags in (5, 1):
            lib.knnsvc_set_option(b"filter_flags", flags)
            for _ in range(3):
                ops.knn_search(qp, pp, K)
            torch.cuda.synchronize()
            lib.knnsvc_filter_timing(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 20
            e0.record()
            for _ in range(n):
                d, i, st = ops.knn_search(qp, pp, K, return_stats=True)
            e1.record(); torch.cuda.synchronize()
            buf = (ctypes.c_float * 256)()
            m = lib.knnsvc_filter_timing_collect(ctypes.cast(buf, ctypes.c_void_p), 256)
            lib.knnsvc_filter_timing(0)
            fl = sorted(buf[j] for j in range(m))
            print(f"{T}x{NP} k={K} rep{rep} flags {flags}: search {e0.elapsed_time(e1)/n:.3f} ms  filter median {fl[m//2]:.3f} min {fl[0]:.3f} ms  "
                  f"n_seg {int(st[3])} units {int(st[4])} grid {int(st[5])}", flush=True)
lib.knnsvc_set_option(b"filter_flags", 1)
