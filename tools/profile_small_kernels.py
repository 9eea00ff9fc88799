"""Launch each non-GEMM kernel once at bench_kernels.py sizes (for an ncu capture)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from knn_svc_b200 import ops, synth
dev = "cuda:0"
g = torch.Generator(device=dev); g.manual_seed(0)
n, T, D = 1_000_000, 100_000, 1024
x = torch.randn((n, D), device=dev, generator=g)
idx = torch.randint(0, n, (T, 4), device=dev, generator=g)
w = torch.softmax(torch.randn((T, 4), device=dev, generator=g), 1)
for _ in range(2):
    ops.prepare_rows(x, check=False)
    ops.gather_mix(x, idx, w)
f0p = torch.rand(n, device=dev, generator=g) * 500 + 80
f0q = torch.rand(T, device=dev, generator=g) * 500 + 80
idx32 = torch.randint(0, n, (T, 32), device=dev, generator=g)
for _ in range(2):
    ops.f0_rerank(f0q, f0p, idx32)
q = torch.from_numpy(synth.ar1_frames(3001, seed=1)).to(dev)
p = torch.from_numpy(synth.ar1_frames(3001, seed=2)).to(dev)
_, nb = ops.knn_search(ops.prepare_rows(q), ops.prepare_rows(p), 4)
f0s = torch.from_numpy(synth.f0_track(3001, seed=3)).to(dev); f0t = torch.from_numpy(synth.f0_track(3001, seed=4)).to(dev)
for _ in range(2):
    ops.concat_cost_reselect(nb, q, p, f0s, f0t)
    ops.weight_fit(nb, p, 0.1)
f0 = torch.from_numpy(np.stack([synth.f0_track(3001, seed=10 + b) for b in range(4)])).to(dev).repeat(16, 1).contiguous()
amp = torch.from_numpy(synth.harmonics_pool(3001, seed=6)).to(dev)[None].repeat(64, 1, 1).contiguous()
for _ in range(2):
    ops.harmonic_bank(f0, amp)
torch.cuda.synchronize()
# ---- SURVEY §8(f) kernels: pool-builder ops, amp_ratio, masked self-search (offline prematch shape)
L, Tl = 25, 3001
layers = torch.randn((L, Tl, D), device=dev, generator=g)
wa = np.random.RandomState(0).rand(L); wb = np.zeros(L); wb[6] = 1.0
audio = torch.randn(320 * Tl + 80, device=dev, generator=g) * 0.1
f0u = torch.from_numpy(synth.f0_track(Tl, seed=7)).to(dev)
spec_pool = torch.rand((n // 4, 200), device=dev, generator=g)
idx4 = torch.randint(0, n // 4, (T, 4), device=dev, generator=g)
for _ in range(2):
    ops.layer_mix(layers, wa, wb)
    spec = ops.stft_magnitude(audio, Tl)
    ops.harmonic_amplitudes(spec, f0u)
    l1 = ops.row_l1(spec_pool)
    ops.amp_ratio(l1[:T], l1, idx4)
lens = torch.full((64,), 512, device=dev)
starts = torch.arange(64, device=dev) * 512
pself = ops.prepare_rows(x[:32768])
for _ in range(2):
    ops.knn_search(pself, pself, 32, mask_lo=torch.repeat_interleave(starts, lens), mask_hi=torch.repeat_interleave(starts + 512, lens))
torch.cuda.synchronize()
