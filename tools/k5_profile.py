"""Phase breakdown of the staged K5 kernel (needs a build with KNNSVC_NVCC_EXTRA=-DKNNSVC_K5_PROFILE)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from knn_svc_b200 import ops, synth
dev = "cuda:0"
T, Np = 3001, 30000
g = torch.Generator(device=dev); g.manual_seed(0)
q = torch.randn((T, 1024), device=dev, generator=g); p = torch.randn((Np, 1024), device=dev, generator=g)
idx = torch.randint(0, Np, (T, 4), device=dev, generator=g)
f0q = torch.rand(T, device=dev, generator=g) * 300 + 100; f0p = torch.rand(Np, device=dev, generator=g) * 300 + 100
for use_f0 in (False, True):
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.concat_cost_reselect(idx, q, p, f0q if use_f0 else None, f0p if use_f0 else None)
        e1.record(); torch.cuda.synchronize()
        print("f0" if use_f0 else "no f0", "ms", e0.elapsed_time(e1), "steps", T - 1, flush=True)
